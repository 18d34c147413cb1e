"""sigmap_b200/paf_eval.py on the committed golden PAF rows (outputs of the unmodified reference,
tests/golden/paf.json) and their simulation truth.  CPU only."""
import numpy as np

from conftest import Golden


def _lines(g, mode):
    # golden rows are stored without the wall-clock mt tag; put one back (position 12, as the
    # reference writes it)
    return ["\t".join(cols[:12] + ["mt:f:1.5"] + cols[12:]) + "\n" for cols in g.paf[mode].values()]


def test_concordance_and_truth_on_golden_rows():
    from sigmap_b200 import paf_eval as E
    g = Golden()
    a = E.read_paf(_lines(g, "default"))
    assert len(a) == len(g.paf["default"]) and all(r.mapped for r in a.values())
    same = E.concordance(a, E.read_paf(_lines(g, "default")))
    assert same["fraction"] == 1.0 and same["identical_rows"] == len(a) and not same["discordant"]
    # full-read mapping of the same reads: same loci, ends may move -> concordant on contig/strand
    full = E.read_paf(_lines(g, "full"))
    loose = E.concordance(a, full, tol=100000)
    assert loose["fraction"] == 1.0
    # a shifted / flipped / unmapped row is discordant
    name = sorted(a)[0]
    r = a[name]
    for bad in (r._replace(t_start=r.t_start + 11), r._replace(strand="-" if r.strand == "+" else "+"),
                r._replace(mapped=False)):
        b = dict(a)
        b[name] = bad
        res = E.concordance(a, b)
        assert res["discordant"] == [name] and res["concordant"] == len(a) - 1
    assert E.concordant(r, r._replace(t_start=r.t_start + 10, t_end=r.t_end - 10))
    # truth: every golden read maps over its simulated origin
    names = [str(n) for n in g["read_names"]]
    lens = g["contig_len"]
    truth = {n: (f"contig_{int(t[0])}", int(t[1]), int(t[2]), "+" if int(t[3]) else "-")
             for n, t in zip(names, g["truth"])}
    assert len(lens) >= 1
    res = E.truth_eval(a, truth)
    assert res["tp"] == len(a) and res["fp"] == res["fn"] == res["tn"] == 0
    assert res["precision"] == res["recall"] == res["f1"] == 1.0
    assert res["mean_time_per_read"] == 1.5 and res["mean_time_per_chunk"] <= 1.5
    # moved origin -> fp; unmapped row -> fn; unmapped without origin -> tn
    moved = dict(truth)
    moved[name] = (truth[name][0], truth[name][2] + 5000, truth[name][2] + 7000, truth[name][3])
    assert E.truth_eval(a, moved)["fp"] == 1
    un = E.parse_line(f"{name}\t{r.read_len}\t*\t*\t*\t*\t*\t*\t*\t*\t*\t61\tci:i:3\n")
    assert not un.mapped and un.mapq == 61 and un.tags["ci"] == 3
    assert E.classify(un, truth[name]) == "fn" and E.classify(un, None) == "tn"


def test_cli_entry_points(tmp_path, capsys):
    from sigmap_b200 import paf_eval as E
    g = Golden()
    p = tmp_path / "a.paf"
    p.write_text("".join(_lines(g, "default")))
    t = tmp_path / "truth.tsv"
    t.write_text("".join(f"{n}\tcontig_{int(x[0])}\t{int(x[1])}\t{int(x[2])}\t{'+' if int(x[3]) else '-'}\n"
                         for n, x in zip([str(n) for n in g["read_names"]], g["truth"])))
    assert E.main(["concordance", str(p), str(p)]) == 0
    assert "100.000 %" in capsys.readouterr().out
    assert E.main(["truth", str(p), str(t)]) == 0
    assert "Sigmap precision: 1.0" in capsys.readouterr().out


def test_rows_sequence_over_a_ctypes_array():
    """Mapping calls hand their rows back as a sequence over the ctypes array the library filled
    (no per-row Python work inside the timed call): length, indexing, slices, iteration."""
    import ctypes as C
    from sigmap_b200 import _ffi as F
    from sigmap_b200.mapper import Rows
    arr = (F.Mapping * 5)()
    for i in range(5):
        arr[i].read_len = 100 + i
    rows = Rows(arr, 4)  # the array may be longer than the read set (never shorter)
    assert len(rows) == 4 and rows[0].read_len == 100 and rows[-1].read_len == 103
    assert [m.read_len for m in rows] == [100, 101, 102, 103]
    assert [m.read_len for m in rows[1:3]] == [101, 102]
    import pytest
    with pytest.raises(IndexError):
        rows[4]
    assert len(Rows((F.Mapping * 1)(), 0)) == 0 and list(Rows((F.Mapping * 1)(), 0)) == []
    assert bytes(rows[2]) == bytes(arr[2]) and C.sizeof(rows[0]) == C.sizeof(F.Mapping)
