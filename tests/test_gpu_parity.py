"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
(oracle/sigmap_oracle.c) and -- where oracle/_ref was built -- the unmodified reference.

Bar: bit-exact for events, hit sets (index + d2 bits), chains, anchors and PAF rows."""
import numpy as np
import pytest

from conftest import bits, paf_cols, same_chains

pytestmark = pytest.mark.gpu

DIG, OFF, RNG = 8192.0, 10.0, 1437.976685


def test_raw_to_pa_filter_compaction(mapper, port, small):
    raw = small.reads.read(0)[:30000].copy()
    # force every branch of the (30, 200) pA filter: values far outside, and at the edges
    rngs = np.random.default_rng(1)
    idx = rngs.choice(len(raw), 600, replace=False)
    raw[idx[:200]] = 3000      # > 200 pA
    raw[idx[200:400]] = -500   # < 30 pA
    raw[idx[400:]] = rngs.integers(150, 1200, 200)
    for edge in (160, 161, 162, 1128, 1129, 1130):
        raw[edge] = edge
    got = mapper.raw_to_pa(raw, DIG, OFF, RNG)
    exp = port.raw_to_pa(raw, DIG, OFF, RNG)
    assert len(exp) < len(raw)
    assert got.shape == exp.shape and np.array_equal(bits(got), bits(exp))
    # nothing kept / everything kept / empty
    assert len(mapper.raw_to_pa(np.full(100, 30000, np.int16), DIG, OFF, RNG)) == 0
    assert len(mapper.raw_to_pa(np.zeros(0, np.int16), DIG, OFF, RNG)) == 0


def test_detect_events_tstat_peaks_means(mapper, port, small):
    for r in range(4):
        pa = small.pa(port, r)
        x = pa[:4000]
        got, exp = mapper.detect_events(x), port.detect_events(x)
        assert np.array_equal(bits(got["tstat1"]), bits(exp["tstat1"]))
        assert np.array_equal(bits(got["tstat2"]), bits(exp["tstat2"]))
        assert np.array_equal(got["peaks"].astype(np.uint64), exp["peaks"])
        assert np.array_equal(bits(got["means"]), bits(exp["means"]))


def test_generate_events_batch(mapper, port, ref, small):
    chunks = []
    for r in range(small.reads.n):
        pa = small.pa(port, r)
        for c in range(min(len(pa) // 4000, 4)):
            chunks.append(pa[c * 4000:(c + 1) * 4000])
    got = mapper.GenerateEvents(np.stack(chunks))
    assert len(got) == len(chunks) > 100
    for x, g in zip(chunks, got):
        e = port.generate_events(x)
        assert g.shape == e.shape and np.array_equal(bits(g), bits(e))
    if ref is not None:  # the reference itself on a subset
        for x, g in list(zip(chunks, got))[:40]:
            e = ref.generate_events(x)
            assert g.shape == e.shape and np.array_equal(bits(g), bits(e))


def test_generate_events_degenerate_chunks(mapper, port):
    flat = np.full(4000, 90.0, np.float32)        # no peaks at all -> no events
    ramp = np.linspace(60, 120, 4000).astype(np.float32)
    rngs = np.random.default_rng(3)
    noise = (90 + 12 * rngs.standard_normal(4000)).astype(np.float32)
    steps = np.repeat(rngs.uniform(60, 120, 400), 10).astype(np.float32)
    got = mapper.GenerateEvents(np.stack([flat, ramp, noise, steps]))
    for x, g in zip((flat, ramp, noise, steps), got):
        e = port.generate_events(x)
        assert g.shape == e.shape and np.array_equal(bits(g), bits(e))
    assert len(got[0]) == 0


def _queries(port, small, n_reads=6):
    qs = []
    for r in range(n_reads):
        f = port.generate_events(small.pa(port, r)[:4000])
        for p in range(2, len(f) - 5, 2):
            qs.append(f[p:p + 6])
    return np.stack(qs)


def test_radius_search_hit_sets(mapper, port, ref, small):
    q = _queries(port, small)
    off, idx, d2 = mapper.radiusSearch(q)
    assert len(off) == len(q) + 1
    total = 0
    h = ref.index_load(small.prefix) if ref is not None and _has_si(small, ref) else None
    for k in range(len(q)):
        gi, gd = idx[off[k]:off[k + 1]], d2[off[k]:off[k + 1]]
        ei, ed = port.radius_search(small.val, q[k])
        assert np.array_equal(gi, ei), f"query {k}: hit set differs"
        assert np.array_equal(bits(gd), bits(ed))
        total += len(ei)
        if h is not None and k % 7 == 0:
            ri, rd = ref.radius_search(h, q[k])
            o = np.argsort(ri)
            # the reference's KD-tree may drop points within 1e-5 of the radius boundary
            keep = np.abs(gd - np.float32(0.08)) > 1e-5
            rkeep = np.abs(rd[o] - np.float32(0.08)) > 1e-5
            assert np.array_equal(gi[keep], ri[o][rkeep])
    assert total > 1000
    if h is not None:
        ref.index_free(h)


def _has_si(small, ref):
    import os
    if not os.path.exists(small.prefix + ".si"):
        from sigmap_b200.host import MODEL_PATH
        # the reference writes <prefix>.pt and <prefix>.si; build into a sibling prefix and
        # check the .pt it wrote equals ours before using its KD-tree
        r = ref.cli(["-i", "-r", small.fasta, "-p", MODEL_PATH, "-o", small.prefix + "_ref"])
        if r.returncode != 0:
            return False
        a = open(small.prefix + ".pt", "rb").read()
        b = open(small.prefix + "_ref.pt", "rb").read()
        assert a == b, "host index builder and reference disagree on the .pt"
        os.replace(small.prefix + "_ref.si", small.prefix + ".si")
    return True


def test_radius_search_edge_queries(mapper, port, small):
    far = np.full((1, 6), 9.0, np.float32)         # no hits
    on_point = small.val[1000:1006][None, :]      # distance 0 to itself
    dense = np.zeros((1, 6), np.float32)
    off, idx, d2 = mapper.radiusSearch(np.concatenate([far, on_point, dense]), radius=0.3)
    assert off[1] == 0
    for k, q in enumerate((far[0], on_point[0], dense[0])):
        ei, ed = port.radius_search(small.val, q, radius=0.3)
        assert np.array_equal(idx[off[k]:off[k + 1]], ei)
        assert np.array_equal(bits(d2[off[k]:off[k + 1]]), bits(ed))
    assert 1000 in idx[off[1]:off[2]]
    # NaN / infinite coordinates (a chunk whose event means are all equal gives 0/0 z-scores):
    # no hits, like nanoflann (NaN < radius is false), and no walk through the whole index
    weird = np.zeros((4, 6), np.float32)
    weird[0, 2] = np.nan
    weird[1, :] = np.nan
    weird[2, 0] = np.inf
    weird[3, 5] = -np.inf
    for name, value in (("search", "lean"), ("search", "general")):
        mapper.set_option(name, value)
        try:
            off, idx, d2 = mapper.radiusSearch(np.concatenate([weird, on_point]), radius=0.08)
        finally:
            mapper.set_option("search", "lean")
        assert list(off[:5]) == [0, 0, 0, 0, 0] and 1000 in idx[off[4]:off[5]]


def test_generate_chains_multi_chunk_state(mapper, port, small):
    n_reads, n_chunks = 12, 4
    feats = {}
    for r in range(n_reads):
        pa = small.pa(port, r)
        feats[r] = [port.generate_events(pa[c * 4000:(c + 1) * 4000])
                    for c in range(min(len(pa) // 4000, n_chunks))]
    batch = mapper.ChainBatch(n_reads)
    lists = {r: port.new_chain_list() for r in range(n_reads)}
    offs = {r: 0 for r in range(n_reads)}
    checked = 0
    for c in range(n_chunks):
        slots = [r for r in range(n_reads) if c < len(feats[r])]
        if not slots:
            break
        batch.GenerateChains(slots, [feats[r][c] for r in slots])
        for r in slots:
            f = feats[r][c]
            if len(f) > 50:
                exp = port.generate_chains(small.pos, small.val, f, offs[r], lists[r],
                                           n_targets=small.ref.n)
                offs[r] += len(f)
            else:
                exp = port.chains_py(lists[r])
            got = batch.chains(r)
            assert same_chains(got, exp), f"read {r} chunk {c}"
            checked += 1
        # slots that sat this round out must keep their chains
        for r in range(n_reads):
            if r not in slots:
                assert same_chains(batch.chains(r), port.chains_py(lists[r]))
    assert checked > 30
    for r in lists:
        port.free_chain_list(lists[r])
    batch.close()


def test_streaming_map_rows_vs_oracle(mapper, port, small):
    rows = mapper.map_reads(small.reads)
    lines = mapper.paf_lines(small.reads, rows, small.ref.names)
    n_mapped = 0
    for r in range(small.reads.n):
        m = port.streaming_map(small.pos, small.val, small.ref.n, small.ref.lengths, small.pa(port, r))
        exp = port.format_paf(m, small.reads.names[r], small.ref.names[m.contig],
                              int(small.ref.lengths[m.contig]), 0.0)
        assert paf_cols(lines[r]) == paf_cols(exp), f"read {r}"
        n_mapped += m.mapped
    assert n_mapped > small.reads.n // 2


def test_streaming_map_rows_vs_reference_cli(mapper, ref, small, tmp_path):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from sigmap_b200.host import MODEL_PATH
    assert _has_si(small, ref)
    out = str(tmp_path / "ref.paf")
    r = ref.cli(["-m", "-r", small.fasta, "-p", MODEL_PATH, "-x", small.prefix, "-s", small.sigdir,
                 "-o", out, "-t", "4"])
    assert r.returncode == 0, r.stderr[-500:]
    exp = {l.split("\t")[0]: paf_cols(l) for l in open(out)}
    rows = mapper.map_reads(small.reads)
    lines = mapper.paf_lines(small.reads, rows, small.ref.names)
    assert len(exp) == small.reads.n
    for name, line in zip(small.reads.names, lines):
        assert paf_cols(line) == exp[name], name


def test_full_read_mode_and_stats(mapper, port, small):
    from sigmap_b200.mapper import full_read_params
    mapper.stats_reset()
    rows = mapper.map_reads(small.reads, full_read_params())
    st = mapper.stats()
    # every chunk of every read is consumed
    for r, m in enumerate(rows):
        n_chunks = len(small.pa(port, r)) // 4000
        assert m.chunks == max(n_chunks, 1)
    assert st["samples"] == sum((len(small.pa(port, r)) // 4000) * 4000 for r in range(small.reads.n))
    assert st["anchors"] >= st["hits"] > 0 and st["queries"] > 0
    # round trip: the mapped position must contain the simulated origin
    ok = 0
    for r, m in enumerate(rows):
        contig, start, end, plus = (int(v) for v in small.reads.truth[r])
        if m.mapped and m.contig == contig and m.strand_plus == plus:
            if m.t_start < end + 50 and m.t_start + m.frag_len > start - 50:
                ok += 1
    assert ok >= 0.9 * small.reads.n


def test_small_batches_give_identical_rows(mapper, small):
    base = mapper.map_reads(small.reads)
    mapper.set_limits(max_batch_chunks=7)
    try:
        again = mapper.map_reads(small.reads)
    finally:
        mapper.set_limits(max_batch_chunks=32768)
    a = mapper.paf_lines(small.reads, base, small.ref.names)
    b = mapper.paf_lines(small.reads, again, small.ref.names)
    assert a == b


def test_streaming_rounds_match_offline(mapper, small):
    """Feed raw samples channel by channel in uneven slices; decisions must equal the offline
    mapping of the same reads (chunk boundaries live on the filtered stream, H9)."""
    n = 16
    offline = mapper.map_reads(small.reads)
    mapper.stream_open(n)
    for ch in range(n):
        mapper.stream_begin_read(ch, DIG, RNG, OFF)
    cursor = [0] * n
    done = {}
    rngs = np.random.default_rng(5)
    for _ in range(200):
        chans, slices = [], []
        for ch in range(n):
            if ch in done:
                continue
            raw = small.reads.read(ch)
            if cursor[ch] >= len(raw):
                continue
            take = int(rngs.integers(1500, 6000))
            chans.append(ch)
            slices.append(raw[cursor[ch]:cursor[ch] + take])
            cursor[ch] += take
        if not chans:
            break
        dec, maps = mapper.stream_round(chans, slices)
        for ch, d, m in zip(chans, dec, maps):
            if d:
                done[ch] = m
    mapper.stream_close()
    for ch in range(n):
        off = offline[ch]
        if ch in done:
            m = done[ch]
            assert (m.mapped, m.contig, m.strand_plus, m.t_start, m.frag_len, m.chunks) == \
                   (off.mapped, off.contig, off.strand_plus, off.t_start, off.frag_len, off.chunks)


def test_streaming_filter_boundaries_match_offline(mapper, host, small):
    """Samples outside (30, 200) pA -- including the raw values either side of both thresholds --
    are dropped by the read-until path (host interval filter) exactly as by the offline K1 kernel."""
    n = 8
    rng = np.random.default_rng(11)
    raws, off_ = [], [0]
    edge = np.array([159, 160, 161, 162, 1128, 1129, 1130, 1131, -32768, 32767, 0, -11], np.int16)
    for r in range(n):
        raw = small.reads.read(r).copy()
        where = rng.choice(len(raw), size=len(raw) // 200, replace=False)
        raw[where] = rng.choice(edge, size=len(where))
        raws.append(raw)
        off_.append(off_[-1] + len(raw))
    spiked = host.ReadSet([f"s{r}" for r in range(n)], np.concatenate(raws), np.array(off_, np.uint64),
                          host.DIGITISATION, host.RANGE, host.OFFSET, small.reads.truth[:n])
    offline = mapper.map_reads(spiked)
    kept = [len(mapper.raw_to_pa(raws[r], DIG, OFF, RNG)) for r in range(n)]
    assert all(k < len(raws[r]) for r, k in enumerate(kept))  # the filter really dropped samples
    mapper.stream_open(n)
    for ch in range(n):
        mapper.stream_begin_read(ch, DIG, RNG, OFF)
    cursor, done = [0] * n, {}
    for _ in range(200):
        chans, slices = [], []
        for ch in range(n):
            if ch in done or cursor[ch] >= len(raws[ch]):
                continue
            take = int(rng.integers(1500, 6000))
            chans.append(ch)
            slices.append(raws[ch][cursor[ch]:cursor[ch] + take])
            cursor[ch] += take
        if not chans:
            break
        dec, maps = mapper.stream_round(chans, slices)
        for ch, d, m in zip(chans, dec, maps):
            if d:
                done[ch] = m
    mapper.stream_close()
    assert done
    for ch, m in done.items():
        o = offline[ch]
        assert (m.mapped, m.contig, m.strand_plus, m.t_start, m.frag_len, m.chunks) == \
               (o.mapped, o.contig, o.strand_plus, o.t_start, o.frag_len, o.chunks)


def test_event_kernels_thread_and_warp_per_chunk_agree(mapper, port, small, monkeypatch):
    """The two event paths (thread per chunk over transposed global arrays; warp per chunk in
    shared memory for small batches) give bit-identical features and PAF rows."""
    from sigmap_b200.mapper import Mapper
    chunks = []
    for r in range(12):
        pa = small.pa(port, r)
        for c in range(min(len(pa) // 4000, 3)):
            chunks.append(pa[c * 4000:(c + 1) * 4000])
    chunks.append(np.full(4000, 90.0, np.float32))                      # no peaks at all
    chunks.append(np.tile(np.array([60.0, 120.0], np.float32), 2000))   # a peak every sample
    chunks = np.stack(chunks)
    base_feat = mapper.GenerateEvents(chunks)
    base_rows = mapper.paf_lines(small.reads, mapper.map_reads(small.reads), small.ref.names)
    for mode in ("thread", "warp"):
        monkeypatch.setenv("SMB_EVENTS", mode)
        m = Mapper(0)
        try:
            m.set_index(small.pos, small.val)
            m.set_contigs(small.ref.lengths)
            got = m.GenerateEvents(chunks)
            assert len(got) == len(base_feat)
            for g, e in zip(got, base_feat):
                assert g.shape == e.shape and np.array_equal(bits(g), bits(e)), mode
            rows = m.paf_lines(small.reads, m.map_reads(small.reads), small.ref.names)
            assert [paf_cols(l) for l in rows] == [paf_cols(l) for l in base_rows], mode
        finally:
            m.close()
