"""CPU tests (no GPU): the C restatement (oracle/sigmap_oracle.c) against the golden vectors
the UNMODIFIED reference produced (tests/golden/, made by tests/make_golden.py).  This is what
pins the oracle; the -m gpu tests then compare the CUDA path with the pinned oracle and with
the same golden vectors.  Bar: bit-exact."""
import hashlib

import numpy as np

from conftest import bits, paf_cols, same_chains


def _pa(port, golden, host, key):
    if key == "spiked":
        return port.raw_to_pa(golden["spiked_raw"], host.DIGITISATION, host.OFFSET, host.RANGE)
    if key.startswith("sim"):
        return port.raw_to_pa(golden.reads(host).read(int(key[3:])), host.DIGITISATION, host.OFFSET,
                              host.RANGE)
    raw, dig, off, rng = golden.real_read(int(key[4:]))
    return port.raw_to_pa(raw, dig, off, rng)


def test_raw_to_pa_matches_reference(port, golden, host):
    keys = [str(k) for k in golden["pa_keys"]]
    for k, sha in zip(keys, golden["pa_sha256"]):
        got = _pa(port, golden, host, k)
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(sha), k
    sp = _pa(port, golden, host, "spiked")
    assert len(sp) < len(golden["spiked_raw"])          # the filter dropped samples
    assert np.array_equal(bits(sp), bits(golden["pa_spiked"]))
    assert np.array_equal(bits(_pa(port, golden, host, "real0")), bits(golden["pa_real0"]))


def _chunks(port, golden, host):
    out = []
    for kind, r, c in golden["chunk_src"]:
        pa = _pa(port, golden, host, ("sim%d" if kind == 0 else "real%d") % r)
        out.append(pa[int(c) * 4000:(int(c) + 1) * 4000])
    return out


def test_generate_events_matches_reference(port, golden, host):
    chunks = _chunks(port, golden, host)
    assert len(chunks) == len(golden["feat_off"]) - 1 >= 30
    for ci, x in enumerate(chunks):
        got = port.generate_events(x)
        exp = golden.chunk_features(ci)
        assert got.shape == exp.shape and np.array_equal(bits(got), bits(exp)), f"chunk {ci}"


def test_detect_events_matches_reference(port, golden, host):
    chunks = _chunks(port, golden, host)
    for k, ci in enumerate(golden["detect_ids"]):
        d = port.detect_events(chunks[int(ci)])
        assert np.array_equal(bits(d["tstat1"]), bits(golden[f"det{k}_t1"]))
        assert np.array_equal(bits(d["tstat2"]), bits(golden[f"det{k}_t2"]))
        assert np.array_equal(d["peaks"], golden[f"det{k}_peaks"])
        assert np.array_equal(bits(d["means"]), bits(golden[f"det{k}_means"]))


def test_point_cloud_matches_reference_pt(port, golden, host, model, tmp_path):
    g = golden.genome(host)
    pos, val = port.build_point_cloud(g.seqs, model[0])
    n, dim, max_leaf = (int(v) for v in golden["pt_n"])
    assert len(pos) == n
    assert np.array_equal(pos[:64], golden["pt_pos_head"]) and np.array_equal(pos[-64:], golden["pt_pos_tail"])
    assert np.array_equal(bits(val[:64]), bits(golden["pt_val_head"]))
    assert np.array_equal(bits(val[-64:]), bits(golden["pt_val_tail"]))
    # the product's host builder + .pt writer reproduce the reference's file byte for byte
    hp, hv = host.build_point_cloud(g, model[0])
    assert np.array_equal(hp, pos) and np.array_equal(bits(hv), bits(val))
    prefix = str(tmp_path / "idx")
    host.write_pt(prefix, hp, hv, dim, max_leaf)
    assert hashlib.sha256(open(prefix + ".pt", "rb").read()).hexdigest() == str(golden["pt_sha256"])
    rp, rv, rd, rl = host.read_pt(prefix)
    assert np.array_equal(rp, hp) and np.array_equal(bits(rv), bits(hv)) and (rd, rl) == (dim, max_leaf)


def test_radius_search_matches_reference_kdtree(port, golden, host, model):
    g = golden.genome(host)
    pos, val = port.build_point_cloud(g.seqs, model[0])
    total = 0
    for name, radius in (("r008", 0.08), ("r030", 0.30)):
        for k, q in enumerate(golden["queries"]):
            gi, gd = port.radius_search(val, q, radius)
            ei, ed = golden.hits(name, k)
            # brute force is exact; the KD-tree may lose points within 1e-5 of the boundary
            # (north_star: "identical except points within 1e-5 of the radius boundary")
            keep = np.abs(gd - np.float32(radius)) > 1e-5
            ekeep = np.abs(ed - np.float32(radius)) > 1e-5
            assert np.array_equal(gi[keep], ei[ekeep]), f"{name} query {k}"
            assert np.array_equal(bits(gd[keep]), bits(ed[ekeep]))
            total += len(ei)
    assert total > 2000


def test_generate_chains_matches_reference(port, golden, host, model):
    g = golden.genome(host)
    pos, val = port.build_point_cloud(g.seqs, model[0])
    feats = {}
    for ci, (kind, r, c) in enumerate(golden["chunk_src"]):
        if kind == 0:
            feats[(int(r), int(c))] = golden.chunk_features(ci)
    lists, offs, n = {}, {}, 0
    for r, c, exp in golden.chain_states():
        if r not in lists:
            lists[r], offs[r] = port.new_chain_list(), 0
        f = feats[(r, c)]
        got = port.generate_chains(pos, val, f, offs[r], lists[r], n_targets=g.n)
        offs[r] += len(f)
        assert same_chains(got, exp), f"read {r} chunk {c}"
        n += 1
    assert n >= 25
    for cl in lists.values():
        port.free_chain_list(cl)


def test_streaming_map_matches_reference_cli(port, golden, host, model):
    g = golden.genome(host)
    reads = golden.reads(host)
    pos, val = port.build_point_cloud(g.seqs, model[0])
    full = port.default_params()
    full.max_num_chunks, full.stop_ratio, full.stop_mean_ratio, full.stop_min_anchors = \
        100000, 1e30, 1e30, 2000000000
    mapped = 0
    for mode, prm in (("default", None), ("full", full)):
        for r, name in enumerate(reads.names):
            pa = port.raw_to_pa(reads.read(r), host.DIGITISATION, host.OFFSET, host.RANGE)
            m = port.streaming_map(pos, val, g.n, g.lengths, pa, prm)
            line = port.format_paf(m, name, g.names[m.contig], int(g.lengths[m.contig]), 0.0)
            assert paf_cols(line) == golden.paf[mode][name], f"{mode} {name}"
            mapped += m.mapped
    assert mapped >= len(reads.names)  # at least half of the 2x10 rows are mapped rows
