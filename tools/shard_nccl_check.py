"""Contig-sharded mapping over NCCL, checked against the unsharded path on the same box.
Run under torchrun (one process per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      --master-port 29517 tools/shard_nccl_check.py
Every rank maps every read against its own contigs; the library's NCCL collectives (sb_exchange.cuh)
must make every rank's rows equal, bit for bit, to the rows of an unsharded mapper built on the
same GPU.  Prints one line per rank and exits non-zero on any difference."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sigmap_b200 import host as H, shard
    from sigmap_b200.mapper import Mapper, default_params, full_read_params
    n_reads = int(os.environ.get("CHECK_READS", "400"))
    model = H.load_pore_model()
    ref = H.sim_reference(11, [900000, 400000, 1200000, 300000, 700000, 550000, 250000, 800000])
    pos, val = H.build_point_cloud(ref, model[0])
    reads = H.sim_reads(12, ref, n_reads, noise=1.3, model=model)
    whole = Mapper(local)
    whole.set_index(pos, val)
    whole.set_contigs(ref.lengths)
    part = Mapper(local)
    shard.nccl_join(part, dist)
    owner = shard.assign_contigs(ref.lengths, world)
    part.set_index_sharded(pos, val, owner)
    part.set_contigs(ref.lengths)
    bad = 0
    for name, prm in (("default", default_params()), ("full-read", full_read_params())):
        exp = whole.map_reads(reads, prm)
        dist.barrier()
        part.stats_reset()
        t0 = time.time()
        got = part.map_reads(reads, prm)
        dt = time.time() - t0
        diff = sum(bytes(a) != bytes(b) for a, b in zip(exp, got))
        bad += diff
        st = part.stats()
        print(f"rank {rank}/{world} {name}: {diff} of {len(exp)} rows differ from the unsharded run; "
              f"{sum(r.mapped for r in got)} mapped; {st['exchanges']} collectives, {st['hits']} local hits, "
              f"{dt * 1e3:.0f} ms", flush=True)
    # the same with every rank building only its own part of the point cloud (genome-scale path)
    part2 = Mapper(local)
    shard.nccl_join(part2, dist)
    cp = H.build_point_cloud_part(ref, model[0], owner, rank)
    part2.set_index_part(cp, ref.n)
    cp.close()
    part2.set_contigs(ref.lengths)
    exp = whole.map_reads(reads, default_params())
    dist.barrier()
    got = part2.map_reads(reads, default_params())
    diff = sum(bytes(a) != bytes(b) for a, b in zip(exp, got))
    bad += diff
    print(f"rank {rank}/{world} own-part index ({part2.num_points} points in the whole cloud): {diff} rows differ", flush=True)
    part2.close()
    # read-sharded: index built on rank 0, broadcast over NCCL, every rank maps its own reads
    rep = Mapper(local)
    shard.nccl_join(rep, dist)
    if rank == 0:
        rep.set_index(pos, val)
        rep.set_contigs(ref.lengths)
    dist.barrier()
    t0 = time.time()
    rep.broadcast_index(0)
    dt = time.time() - t0
    rep.set_contigs(ref.lengths)
    mine = shard.shard_reads(reads, world, rank)
    lo, hi = shard.block_range(reads.n, world, rank)
    got = rep.map_reads(mine, default_params())
    diff = sum(bytes(a) != bytes(b) for a, b in zip(exp[lo:hi], got))
    bad += diff
    print(f"rank {rank}/{world} broadcast index: {dt * 1e3:.1f} ms, {diff} of {len(got)} rows differ", flush=True)
    rep.close()
    t = torch.tensor([bad], device=f"cuda:{local}")
    dist.all_reduce(t)
    whole.close()
    part.close()
    dist.destroy_process_group()
    if int(t.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("NCCL contig-shard check: PASS", flush=True)


if __name__ == "__main__":
    main()
