"""CPU test of the N>1 path: world_size 2 over gloo.  Each rank takes its block of the golden
reads, "maps" it (with the CPU oracle -- this is a test of the sharding/gather logic, the GPU
mapping itself is covered by the -m gpu tests), and rank 0 must end up with every row, in read
order, equal to the reference CLI's PAF."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, paf_cols


def test_block_range_partitions_exactly():
    from sigmap_b200.shard import block_range
    for n in (0, 1, 7, 10, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [block_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_assign_contigs_balanced_and_deterministic():
    """Contig-sharded index: longest-first bin packing onto the lightest rank."""
    from sigmap_b200.shard import assign_contigs
    rng = np.random.default_rng(5)
    human = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
             138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
             83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
    for lengths, world in ((human, 8), (human, 3), (rng.integers(1000, 10 ** 6, 40).tolist(), 5),
                           ([10, 10, 10], 8), ([7], 1)):
        owner = assign_contigs(lengths, world)
        assert owner.shape == (len(lengths),) and owner.max() < world
        assert np.array_equal(owner, assign_contigs(lengths, world))
        load = np.bincount(owner, weights=np.asarray(lengths, float), minlength=world)
        # LPT bound: no rank above 4/3 of the ideal split unless one contig alone exceeds it
        assert load.max() <= max(max(lengths), 4.0 / 3.0 * sum(lengths) / world + 1)
        if len(lengths) >= world:
            assert (load > 0).all()
    with pytest.raises(ValueError):
        assign_contigs([5, 6], 0)


def test_shard_reads_slices(golden, host):
    from sigmap_b200.shard import shard_reads
    reads = golden.reads(host)
    parts = [shard_reads(reads, 3, r) for r in range(3)]
    assert sum(p.n for p in parts) == reads.n
    k = 0
    for p in parts:
        for i in range(p.n):
            assert p.names[i] == reads.names[k] and np.array_equal(p.read(i), reads.read(k))
            k += 1


def _worker(rank, world, port_no, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import Golden
    from oracle.oracle import Port
    from sigmap_b200 import _ffi, host as H
    from sigmap_b200.shard import gather_rows, reduce_scalar, shard_reads
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    g = Golden()
    genome, reads = g.genome(H), g.reads(H)
    port = Port()
    pos, val = port.build_point_cloud(genome.seqs, H.load_pore_model()[0])
    mine = shard_reads(reads, world, rank)
    rows = []
    for r in range(mine.n):
        pa = port.raw_to_pa(mine.read(r), H.DIGITISATION, H.OFFSET, H.RANGE)
        m = port.streaming_map(pos, val, genome.n, genome.lengths, pa)
        row = _ffi.Mapping()
        for f, _ in _ffi.Mapping._fields_:
            if hasattr(m, f):
                setattr(row, f, getattr(m, f))
        rows.append(row)
    allrows = gather_rows(rows, _ffi.Mapping, dist)
    total = reduce_scalar(mine.n, "sum", dist)
    slowest = reduce_scalar(float(rank + 1), "max", dist)
    if rank == 0:
        assert total == reads.n and slowest == world
        lines = [H.format_paf(m, n, genome.names[m.contig] if m.mapped else "",
                              int(genome.lengths[m.contig]) if m.mapped else 0, 0.0)
                 for n, m in zip(reads.names, allrows)]
        with open(out_path, "w") as f:
            f.writelines(l if l.endswith("\n") else l + "\n" for l in lines)
    else:
        assert allrows is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_rows_in_read_order(golden, tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    out = str(tmp_path / "gathered.paf")
    mp.spawn(_worker, args=(2, port_no, out), nprocs=2, join=True)
    lines = open(out).read().splitlines()
    assert len(lines) == len(golden.paf["default"])
    for line in lines:
        cols = paf_cols(line)
        assert cols == golden.paf["default"][cols[0]]
