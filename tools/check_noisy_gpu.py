#!/usr/bin/env python
"""CUDA path vs the C oracle on inputs harder than the test suite's: three contigs, noise 1.6,
a third of the reads carrying samples outside the (30, 200) pA window; default and full-read
parameters.  Prints the number of differing PAF rows; exit code 1 if any.  (Round-2 candidate for
a `-m gpu` test once it has been seen to pass on hardware.)"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import Dataset, paf_cols
    from oracle.oracle import Port
    from sigmap_b200 import host as H
    from sigmap_b200.mapper import Mapper, default_params, full_read_params
    model = H.load_pore_model()
    ds = Dataset(H, model, tempfile.mkdtemp(prefix="noisy_"), [150000, 80000, 120000], 36, seed=99, noise=1.6,
                 min_bases=1500, max_bases=7000)
    rng = np.random.default_rng(3)
    raw = ds.reads.raw.copy()
    edge = np.array([159, 160, 161, 162, 1128, 1129, 1130, 1131, -32768, 32767, 0], np.int16)
    for r in range(0, ds.reads.n, 3):
        a, b = int(ds.reads.read_off[r]), int(ds.reads.read_off[r + 1])
        where = a + rng.choice(b - a, size=(b - a) // 150, replace=False)
        raw[where] = rng.choice(edge, size=len(where))
    reads = H.ReadSet(ds.reads.names, raw, ds.reads.read_off, H.DIGITISATION, H.RANGE, H.OFFSET, ds.reads.truth)
    port = Port()
    m = Mapper(0)
    m.set_index(ds.pos, ds.val)
    m.set_contigs(ds.ref.lengths)
    bad = 0
    full = port.default_params()
    full.max_num_chunks, full.stop_ratio, full.stop_mean_ratio, full.stop_min_anchors = 100000, 1e30, 1e30, 2000000000
    for mode, gp, op in (("default", default_params(), None), ("full", full_read_params(), full)):
        rows = m.map_reads(reads, gp)
        lines = m.paf_lines(reads, rows, ds.ref.names)
        for i, name in enumerate(reads.names):
            pa = port.raw_to_pa(reads.read(i), H.DIGITISATION, H.OFFSET, H.RANGE)
            e = port.streaming_map(ds.pos, ds.val, ds.ref.n, ds.ref.lengths, pa, op)
            exp = port.format_paf(e, name, ds.ref.names[e.contig], int(ds.ref.lengths[e.contig]), 0.0)
            if paf_cols(lines[i]) != paf_cols(exp):
                bad += 1
                print(f"DIFF {mode} {name}\n  gpu {lines[i].strip()}\n  cpu {exp.strip()}")
    m.close()
    print(f"noisy check: {2 * reads.n} rows compared, {bad} differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
