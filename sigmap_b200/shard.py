"""Read sharding across GPUs (mode 1): reads are independent
(sigmap.cc:630-866 touches only its own read and the read-only index), so ranks map disjoint
slices of the read set against a replicated index and the only exchange is the final gather of
the fixed-size result rows.  No data-path collective.

Works on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np


def block_range(n_items, world, rank):
    """Contiguous, balanced slice [lo, hi) of n_items owned by `rank` (first n % world ranks get
    one extra item).  Concatenating the slices in rank order restores the original order."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_reads(reads, world, rank):
    """The ReadSet slice of this rank (views into the same raw buffer; offsets rebased)."""
    from .host import ReadSet
    lo, hi = block_range(reads.n, world, rank)
    s0, s1 = int(reads.read_off[lo]), int(reads.read_off[hi])
    sub = ReadSet(reads.names[lo:hi], reads.raw[s0:s1], reads.read_off[lo:hi + 1] - np.uint64(s0),
                  1.0, 1.0, 0.0, None if reads.truth is None else reads.truth[lo:hi])
    sub.digitisation = reads.digitisation[lo:hi].copy()
    sub.range = reads.range[lo:hi].copy()
    sub.offset = reads.offset[lo:hi].copy()
    return sub


def rows_to_array(rows, struct_type):
    """list of ctypes structs -> uint8 array (n, sizeof)"""
    size = C.sizeof(struct_type)
    out = np.zeros((len(rows), size), np.uint8)
    for i, r in enumerate(rows):
        out[i] = np.frombuffer(bytes(r), np.uint8)
    return out


def array_to_rows(arr, struct_type):
    return [struct_type.from_buffer_copy(arr[i].tobytes()) for i in range(arr.shape[0])]


def gather_rows(rows, struct_type, dist=None, device="cpu"):
    """Every rank contributes its result rows (ctypes structs); rank 0 returns all rows in rank
    order (= read order under block_range sharding), other ranks return None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(rows)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.from_numpy(rows_to_array(rows, struct_type)).to(device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([mine.shape[0]], dtype=torch.int64, device=device))
    counts = [int(c.item()) for c in counts]
    width = C.sizeof(struct_type)
    padded = torch.zeros((max(counts + [1]), width), dtype=torch.uint8, device=device)
    padded[:mine.shape[0]] = mine
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        out.extend(array_to_rows(bufs[r][:counts[r]].cpu().numpy(), struct_type))
    return out


def reduce_scalar(x, op, dist=None, device="cpu"):
    """max / sum of a Python float over ranks (timing is the max over ranks, work the sum)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())
