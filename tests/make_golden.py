#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference).  TEST INFRASTRUCTURE.

The reference ships no golden vectors for the mapping path, so the pins are
outputs of the reference itself on committed inputs:

  golden/stages.npz   inputs (synthetic reference, simulated raw reads, one real R9.4 read
                      slice from slow5lib's test data) and, for each stage, what the
                      reference's own code returned:
                        raw -> pA + (30,200) filter      SignalBatch::AddSignal
                        t-stats / peaks / event means    DetectEvents (event.h:226)
                        per-chunk features               Sigmap::GenerateEvents
                        radius-search hit sets           nanoflann radiusSearch on the .si
                        chains after every chunk         SpatialIndex::GenerateChains
                        .pt checksum                     sigmap -i
  golden/paf.json     PAF rows of `sigmap -m` (default flags, and the full-read flags of
                      SURVEY.md 8d) on the same reads

Run here (needs /root/reference for the build, not at test time):
    python tests/make_golden.py
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

FULL_READ_CLI = ["--max-num-chunks", "100000", "--stop-mapping", "1e30", "--stop-mapping-mean",
                 "1e30", "--min-num-anchors", "2000000000"]
REAL_BLOW5 = "/root/reference/extern/slow5lib/test/data/exp/two_rg/exp_default.blow5"
N_CHUNKS = 3            # chunks per read for the stage vectors
REAL_SAMPLES = 14000    # samples kept of each real read


def chains_flat(chains):
    """list of chain dicts -> (records[n,7] uint32 + score bits, anchors[m,3] uint32)"""
    rec, anc = [], []
    for c in chains:
        rec.append([np.float32(c["score"]).view(np.uint32), c["contig"], c["start"], c["end"],
                    c["n_anchors"], c["mapq"], c["dir"], len(c["anchors"])])
        for t, q, d in c["anchors"]:
            anc.append([t, q, np.float32(d).view(np.uint32)])
    return (np.array(rec, np.uint32).reshape(-1, 8), np.array(anc, np.uint32).reshape(-1, 3))


def main():
    from oracle.oracle import Ref, build
    from sigmap_b200 import host as H
    build()
    assert Ref.available(), "oracle/_ref was not built (is /root/reference mounted?)"
    ref = Ref()
    model = H.load_pore_model()
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp(prefix="golden_")
    out = {}
    try:
        # ---------------- inputs
        genome = H.sim_reference(11, [60000, 45000])
        reads = H.sim_reads(12, genome, 10, min_bases=1400, max_bases=2600, noise=1.0, model=model)
        real = H.ReadSet.read_blow5(REAL_BLOW5)
        out["contig_len"] = genome.lengths
        out["contig_seq"] = np.frombuffer(b"".join(genome.seqs), np.uint8)
        out["raw"], out["read_off"] = reads.raw, reads.read_off
        out["read_names"] = np.array(reads.names)
        out["truth"] = reads.truth
        real_raw = [real.read(r)[:REAL_SAMPLES].copy() for r in range(real.n)]
        out["real_raw"] = np.concatenate(real_raw)
        out["real_off"] = np.cumsum([0] + [len(x) for x in real_raw]).astype(np.uint64)
        out["real_dig"], out["real_range"], out["real_offset"] = real.digitisation, real.range, real.offset

        # ---------------- index: the reference builds .pt/.si from the FASTA
        fasta = os.path.join(work, "ref.fa")
        genome.write_fasta(fasta)
        prefix = os.path.join(work, "idx")
        r = ref.cli(["-i", "-r", fasta, "-p", H.MODEL_PATH, "-o", prefix])
        assert r.returncode == 0, r.stderr
        pt = open(prefix + ".pt", "rb").read()
        out["pt_sha256"] = np.array(hashlib.sha256(pt).hexdigest())
        pos, val, dim, max_leaf = H.read_pt(prefix)
        out["pt_n"] = np.array([len(pos), dim, max_leaf], np.uint64)
        out["pt_pos_head"], out["pt_val_head"] = pos[:64], val[:64]
        out["pt_pos_tail"], out["pt_val_tail"] = pos[-64:], val[-64:]
        h = ref.index_load(prefix)

        # ---------------- raw -> pA (synthetic, real, and a spiked copy that exercises the filter)
        spiked = reads.read(0).copy()
        rng = np.random.default_rng(5)
        idx = rng.choice(len(spiked), 300, replace=False)
        spiked[idx[:100]] = 3000
        spiked[idx[100:200]] = -500
        spiked[idx[200:]] = rng.integers(150, 1200, 100)
        for edge in (160, 161, 162, 1128, 1129, 1130):
            spiked[edge] = edge
        out["spiked_raw"] = spiked
        pa_sets = {"spiked": ref.raw_to_pa(spiked, H.DIGITISATION, H.OFFSET, H.RANGE)}
        for r_ in range(reads.n):
            pa_sets[f"sim{r_}"] = ref.raw_to_pa(reads.read(r_), H.DIGITISATION, H.OFFSET, H.RANGE)
        for r_ in range(real.n):
            pa_sets[f"real{r_}"] = ref.raw_to_pa(real_raw[r_], float(real.digitisation[r_]),
                                                 float(real.offset[r_]), float(real.range[r_]))
        out["pa_spiked"] = pa_sets["spiked"]
        out["pa_real0"] = pa_sets["real0"]
        out["pa_kept_len"] = np.array([len(pa_sets[f"sim{r_}"]) for r_ in range(reads.n)] +
                                      [len(pa_sets[f"real{r_}"]) for r_ in range(real.n)], np.uint64)
        out["pa_sha256"] = np.array([hashlib.sha256(pa_sets[k].tobytes()).hexdigest()
                                     for k in sorted(pa_sets)])
        out["pa_keys"] = np.array(sorted(pa_sets))

        # ---------------- events: chunks of the simulated reads + of the real reads
        chunk_src, chunks = [], []
        for r_ in range(reads.n):
            pa = pa_sets[f"sim{r_}"]
            for c in range(min(len(pa) // 4000, N_CHUNKS)):
                chunk_src.append((0, r_, c))
                chunks.append(pa[c * 4000:(c + 1) * 4000])
        for r_ in range(real.n):
            pa = pa_sets[f"real{r_}"]
            for c in range(min(len(pa) // 4000, N_CHUNKS)):
                chunk_src.append((1, r_, c))
                chunks.append(pa[c * 4000:(c + 1) * 4000])
        out["chunk_src"] = np.array(chunk_src, np.uint32)
        feats = [ref.generate_events(x) for x in chunks]
        out["feat"] = np.concatenate(feats)
        out["feat_off"] = np.cumsum([0] + [len(f) for f in feats]).astype(np.uint64)
        det_ids = [0, 1, len(chunks) - 1]          # two simulated chunks and a real one
        out["detect_ids"] = np.array(det_ids, np.uint32)
        for k, ci in enumerate(det_ids):
            d = ref.detect_events(chunks[ci])
            out[f"det{k}_t1"], out[f"det{k}_t2"] = d["tstat1"], d["tstat2"]
            out[f"det{k}_peaks"], out[f"det{k}_means"] = d["peaks"], d["means"]

        # ---------------- radius search hit sets (KD-tree order -> sorted by index)
        q_list = []
        for ci in range(0, len(chunks), 2):
            f = feats[ci]
            for p in range(2, len(f) - 5, 16):
                q_list.append(f[p:p + 6])
        q_list = np.stack(q_list[:160])
        out["queries"] = q_list
        for name, radius in (("r008", 0.08), ("r030", 0.30)):
            off, ids, d2s = [0], [], []
            for q in q_list:
                i, d = ref.radius_search(h, q, radius)
                o = np.argsort(i, kind="stable")
                ids.append(i[o])
                d2s.append(d[o])
                off.append(off[-1] + len(i))
            out[f"hits_{name}_off"] = np.array(off, np.uint64)
            out[f"hits_{name}_idx"] = np.concatenate(ids)
            out[f"hits_{name}_d2"] = np.concatenate(d2s)

        # ---------------- chains after every chunk, state carried like StreamingMap does
        ch_rec, ch_anc, ch_key = [], [], []
        ci = 0
        for r_ in range(reads.n):
            n_c = sum(1 for s in chunk_src if s[0] == 0 and s[1] == r_)
            st = ref.chain_state_new()
            offset = 0
            for c in range(n_c):
                f = feats[ci]
                if len(f) > 50:   # sigmap.cc:660
                    chains = ref.generate_chains(h, st, f, offset, n_targets=genome.n)
                    offset += len(f)
                    rec, anc = chains_flat(chains)
                    ch_key.append((r_, c, len(rec), len(anc)))
                    ch_rec.append(rec)
                    ch_anc.append(anc)
                ci += 1
            ref.chain_state_free(st)
        out["chain_key"] = np.array(ch_key, np.uint32)
        out["chain_rec"] = np.concatenate(ch_rec)
        out["chain_anc"] = np.concatenate(ch_anc)
        ref.index_free(h)

        # ---------------- PAF rows of the reference CLI
        sig = os.path.join(work, "sig")
        os.makedirs(sig)
        reads.write_blow5(os.path.join(sig, "reads.blow5"))
        paf = {}
        for mode, extra in (("default", []), ("full", FULL_READ_CLI)):
            o = os.path.join(work, mode + ".paf")
            r = ref.cli(["-m", "-r", fasta, "-p", H.MODEL_PATH, "-x", prefix, "-s", sig, "-o", o,
                         "-t", "2"] + extra)
            assert r.returncode == 0, r.stderr
            rows = {}
            for line in open(o):
                cols = [c for c in line.rstrip("\n").split("\t") if not c.startswith("mt:f:")]
                rows[cols[0]] = cols
            paf[mode] = rows
        json.dump(paf, open(os.path.join(GOLD, "paf.json"), "w"), indent=0, sort_keys=True)
        np.savez_compressed(os.path.join(GOLD, "stages.npz"), **out)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    sz = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"golden vectors written to {GOLD}: {sz / 1024:.0f} KiB, {len(chunks)} chunks, "
          f"{len(q_list)} queries, {len(ch_key)} chain states")


if __name__ == "__main__":
    main()
