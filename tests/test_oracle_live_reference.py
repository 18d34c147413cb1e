"""The C restatement (oracle/sigmap_oracle.c) against the UNMODIFIED reference run live on fresh
inputs -- beyond the committed golden vectors: several contigs, noisier signal (more candidate
chains, more chunks per read), reads with samples outside (30, 200) pA, and non-default mapping
parameters.  CPU only; skipped where oracle/_ref was not built (the GPU box has it prebuilt, the
reference sources exist only in the build container)."""
import os

import numpy as np
import pytest

from conftest import Dataset, paf_cols


@pytest.fixture(scope="module")
def noisy(host, model, tmp_path_factory):
    ds = Dataset(host, model, tmp_path_factory.mktemp("noisy"), [150000, 80000, 120000], 24, seed=99,
                 noise=1.6, min_bases=1500, max_bases=7000)
    # a third of the reads get samples outside the pA window, including the raw values either
    # side of both thresholds; the BLOW5 the reference reads is rewritten with them
    rng = np.random.default_rng(3)
    raw = ds.reads.raw.copy()
    edge = np.array([159, 160, 161, 162, 1128, 1129, 1130, 1131, -32768, 32767, 0], np.int16)
    for r in range(0, ds.reads.n, 3):
        a, b = int(ds.reads.read_off[r]), int(ds.reads.read_off[r + 1])
        where = a + rng.choice(b - a, size=(b - a) // 150, replace=False)
        raw[where] = rng.choice(edge, size=len(where))
    ds.reads = host.ReadSet(ds.reads.names, raw, ds.reads.read_off, host.DIGITISATION, host.RANGE,
                            host.OFFSET, ds.reads.truth)
    ds.reads.write_blow5(os.path.join(ds.sigdir, "reads.blow5"))
    return ds


@pytest.mark.parametrize("mode", ["default", "tuned"])
def test_restatement_rows_equal_live_reference_rows(port, ref, host, noisy, tmp_path, mode):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from sigmap_b200.host import MODEL_PATH
    ds = noisy
    idx = str(tmp_path / "refidx")
    r = ref.cli(["-i", "-r", ds.fasta, "-p", MODEL_PATH, "-o", idx])
    assert r.returncode == 0, r.stderr[-400:]
    assert open(idx + ".pt", "rb").read() == open(ds.prefix + ".pt", "rb").read()
    extra, prm = [], port.default_params()
    if mode == "tuned":
        extra = ["--step-size", "3", "--search-radius", "0.1", "--max-num-chunks", "6",
                 "--stop-mapping", "1.8", "--stop-mapping-mean", "7", "--min-num-anchors", "12"]
        prm.step, prm.search_radius, prm.max_num_chunks = 3, 0.1, 6
        prm.stop_ratio, prm.stop_mean_ratio, prm.stop_min_anchors = 1.8, 7.0, 12
    out = str(tmp_path / "ref.paf")
    r = ref.cli(["-m", "-r", ds.fasta, "-p", MODEL_PATH, "-x", idx, "-s", ds.sigdir, "-o", out, "-t", "4"] + extra)
    assert r.returncode == 0, r.stderr[-400:]
    exp = {l.split("\t")[0]: paf_cols(l) for l in open(out)}
    assert len(exp) == ds.reads.n
    mapped = chunks = 0
    for i, name in enumerate(ds.reads.names):
        pa = port.raw_to_pa(ds.reads.read(i), host.DIGITISATION, host.OFFSET, host.RANGE)
        m = port.streaming_map(ds.pos, ds.val, ds.ref.n, ds.ref.lengths, pa, prm)
        line = port.format_paf(m, name, ds.ref.names[m.contig], int(ds.ref.lengths[m.contig]), 0.0)
        assert paf_cols(line) == exp[name], f"{mode} {name}"
        mapped += m.mapped
        chunks += int(exp[name][12].split(":")[2])
    assert mapped >= ds.reads.n // 2
    if mode == "default":
        assert chunks > ds.reads.n  # the noise makes some reads need more than one chunk


def test_reference_loads_and_uses_our_si(host, model, ref, port, tmp_path):
    """`sigmap -i` of this repo writes <prefix>.pt and <prefix>.si; the UNMODIFIED reference must be
    able to load them (SpatialIndex::Load, spatial_index.cc:132-163) and map with them exactly as
    with an index it built itself: same radius-search hit sets, same PAF."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from conftest import Dataset, paf_cols
    from sigmap_b200.host import MODEL_PATH
    ds = Dataset(host, model, tmp_path, [180000, 90000], 40, seed=17)
    host.write_si(ds.prefix, ds.val)                      # next to the .pt the fixture wrote
    own = str(tmp_path / "own")
    r = ref.cli(["-i", "-r", ds.fasta, "-p", MODEL_PATH, "-o", own])
    assert r.returncode == 0, r.stderr[-300:]
    assert open(own + ".pt", "rb").read() == open(ds.prefix + ".pt", "rb").read()
    # stage level: the reference's radiusSearch over OUR tree = brute force
    h = ref.index_load(ds.prefix)
    n_hits = 0
    for rd in range(3):
        f = port.generate_events(ds.pa(port, rd)[:4000])
        for p in range(2, len(f) - 5, 16):
            q = f[p:p + 6]
            ri, rdist = ref.radius_search(h, q)
            ei, ed = port.radius_search(ds.val, q)
            o = np.argsort(ri)
            keep = np.abs(ed - np.float32(0.08)) > 1e-5
            rkeep = np.abs(rdist[o] - np.float32(0.08)) > 1e-5
            assert np.array_equal(ri[o][rkeep], ei[keep])
            n_hits += len(ei)
    ref.index_free(h)
    assert n_hits > 100
    # whole path: the reference CLI with our index files vs with its own
    outs = []
    for prefix in (ds.prefix, own):
        out = str(tmp_path / (os.path.basename(prefix) + ".paf"))
        r = ref.cli(["-m", "-r", ds.fasta, "-p", MODEL_PATH, "-x", prefix, "-s", ds.sigdir, "-o", out, "-t", "4"])
        assert r.returncode == 0, r.stderr[-300:]
        outs.append({l.split("\t")[0]: paf_cols(l) for l in open(out)})
    assert len(outs[0]) == ds.reads.n and outs[0] == outs[1]
