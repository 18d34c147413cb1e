"""Evaluation of Sigmap modified-PAF files (host-side tooling next to the mapping path).

Two questions, both keyed by read name (row order differs between runs):

* concordance of two PAF files of the same reads -- the parity criterion of BASELINE.json:
  same mapped/unmapped decision, same contig and strand, target start and end within a
  tolerance (10 bp);
* accuracy against the simulation truth (contig, start, end, strand per read): the TP/FP/FN/TN
  counts, precision, recall, F1 and the time-per-chunk / time-per-read summaries that the
  reference's eval/eval.py prints (eval.py:55-111; it needs PAFs annotated by UNCALLED's
  pafstats, which a simulated truth table replaces here).

    python -m sigmap_b200.paf_eval concordance A.paf B.paf [--tol 10]
    python -m sigmap_b200.paf_eval truth OUT.paf TRUTH.tsv [--slack 50]

TRUTH.tsv: `read_name<TAB>contig_name<TAB>start<TAB>end<TAB>strand(+/-)` per line.
No GPU and no native library needed.
"""
import argparse
import statistics
import sys
from collections import namedtuple

Row = namedtuple("Row", "name read_len mapped q_start q_end strand contig contig_len t_start t_end mapq tags")


def parse_line(line):
    """One modified-PAF row (output_tools.h:200-210,336-354).  Unmapped rows carry `*` in the
    nine alignment columns and mapq 61 (sigmap.cc:860-864)."""
    c = line.rstrip("\n").split("\t")
    if len(c) < 12:
        raise ValueError(f"not a PAF row ({len(c)} columns): {line[:80]!r}")
    tags = {}
    for t in c[12:]:
        k = t.split(":", 2)
        if len(k) == 3:
            tags[k[0]] = int(k[2]) if k[1] == "i" else float(k[2]) if k[1] == "f" else k[2]
    if c[4] == "*":
        return Row(c[0], int(c[1]), False, None, None, None, None, None, None, None, int(c[11]), tags)
    return Row(c[0], int(c[1]), True, int(c[2]), int(c[3]), c[4], c[5], int(c[6]), int(c[7]), int(c[8]),
               int(c[11]), tags)


def read_paf(path_or_lines):
    lines = open(path_or_lines) if isinstance(path_or_lines, str) else path_or_lines
    rows = {}
    for line in lines:
        if line.strip():
            r = parse_line(line)
            rows[r.name] = r
    return rows


def concordant(a, b, tol=10):
    """BASELINE.json's criterion for one read present in both files."""
    if a.mapped != b.mapped:
        return False
    if not a.mapped:
        return True
    return (a.contig == b.contig and a.strand == b.strand and abs(a.t_start - b.t_start) <= tol and
            abs(a.t_end - b.t_end) <= tol)


def concordance(rows_a, rows_b, tol=10):
    names = sorted(set(rows_a) | set(rows_b))
    both = [n for n in names if n in rows_a and n in rows_b]
    ok = [n for n in both if concordant(rows_a[n], rows_b[n], tol)]
    return {
        "reads": len(names), "in_both": len(both), "only_a": sum(n not in rows_b for n in names),
        "only_b": sum(n not in rows_a for n in names), "concordant": len(ok),
        "discordant": sorted(set(both) - set(ok)),
        "fraction": len(ok) / len(names) if names else 1.0,
        "identical_rows": sum(rows_a[n][:11] == rows_b[n][:11] for n in both),
    }


def read_truth(path_or_lines):
    lines = open(path_or_lines) if isinstance(path_or_lines, str) else path_or_lines
    truth = {}
    for line in lines:
        c = line.rstrip("\n").split("\t")
        if len(c) >= 5 and not line.startswith("#"):
            truth[c[0]] = (c[1], int(c[2]), int(c[3]), c[4])
    return truth


def classify(row, origin, slack=50):
    """tp: mapped over its origin; fp: mapped elsewhere (or mapped with no origin);
    fn: unmapped although it has an origin; tn: unmapped and no origin."""
    if origin is None:
        return "fp" if row.mapped else "tn"
    if not row.mapped:
        return "fn"
    contig, start, end, strand = origin
    hit = (row.contig == contig and row.strand == strand and row.t_start < end + slack and
           row.t_end > start - slack)
    return "tp" if hit else "fp"


def truth_eval(rows, truth, slack=50):
    counts = {"tp": 0, "fp": 0, "fn": 0, "tn": 0}
    per_chunk, per_read = [], []
    for name, r in rows.items():
        counts[classify(r, truth.get(name), slack)] += 1
        mt = r.tags.get("mt")
        if mt is not None:
            per_read.append(float(mt))
            per_chunk.append(float(mt) / max(int(r.tags.get("ci", 1)), 1))
    tp, fp, fn = counts["tp"], counts["fp"], counts["fn"]
    precision = tp / (tp + fp) if tp + fp else 0.0
    recall = tp / (tp + fn) if tp + fn else 0.0
    out = dict(counts)
    out.update({
        "precision": precision, "recall": recall,
        "f1": 2 * precision * recall / (precision + recall) if precision + recall else 0.0,
        "mean_chunks": statistics.mean(int(r.tags.get("ci", 1)) for r in rows.values()) if rows else 0.0,
        "mean_time_per_chunk": statistics.mean(per_chunk) if per_chunk else None,
        "median_time_per_chunk": statistics.median(per_chunk) if per_chunk else None,
        "mean_time_per_read": statistics.mean(per_read) if per_read else None,
        "median_time_per_read": statistics.median(per_read) if per_read else None,
    })
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("concordance")
    c.add_argument("a")
    c.add_argument("b")
    c.add_argument("--tol", type=int, default=10)
    t = sub.add_parser("truth")
    t.add_argument("paf")
    t.add_argument("truth")
    t.add_argument("--slack", type=int, default=50)
    args = ap.parse_args(argv)
    if args.cmd == "concordance":
        res = concordance(read_paf(args.a), read_paf(args.b), args.tol)
        print(f"reads: {res['reads']} (in both: {res['in_both']}, only A: {res['only_a']}, only B: {res['only_b']})")
        print(f"concordant: {res['concordant']} ({100.0 * res['fraction']:.3f} %), "
              f"identical alignment columns: {res['identical_rows']}")
        for n in res["discordant"][:20]:
            print("discordant:", n)
        return 0 if res["fraction"] >= 0.995 else 1
    res = truth_eval(read_paf(args.paf), read_truth(args.truth), args.slack)
    print("Sigmap TP: %d\nSigmap FP: %d\nSigmap FN: %d\nSigmap TN: %d" % (res["tp"], res["fp"], res["fn"], res["tn"]))
    print("Sigmap precision: %s\nSigmap recall: %s\nSigmap F-1 score: %s" % (res["precision"], res["recall"], res["f1"]))
    for k in ("mean_time_per_chunk", "median_time_per_chunk", "mean_time_per_read", "median_time_per_read"):
        print(k.replace("_", " ").capitalize(), ":", res[k])
    return 0


if __name__ == "__main__":
    sys.exit(main())
