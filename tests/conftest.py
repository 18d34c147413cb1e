"""Shared fixtures.  `-m "not gpu"` tests run on the CPU-only build box; `-m gpu` tests are the
parity tests proper and call the CUDA path through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def host():
    from sigmap_b200 import host as H
    return H


@pytest.fixture(scope="session")
def model(host):
    return host.load_pore_model()


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref), when it was built; else None."""
    from oracle.oracle import Ref
    return Ref() if Ref.available() else None


class Dataset:
    """A small synthetic reference + reads + point cloud, built with the host helpers."""

    def __init__(self, host, model, tmp, contig_lengths, n_reads, seed=7, noise=1.0,
                 min_bases=2000, max_bases=9000):
        self.dir = str(tmp)
        self.ref = host.sim_reference(seed, contig_lengths)
        self.fasta = os.path.join(self.dir, "ref.fa")
        self.ref.write_fasta(self.fasta)
        self.pos, self.val = host.build_point_cloud(self.ref, model[0])
        self.prefix = os.path.join(self.dir, "idx")
        host.write_pt(self.prefix, self.pos, self.val)
        self.reads = host.sim_reads(seed + 4, self.ref, n_reads, noise=noise, min_bases=min_bases,
                                    max_bases=max_bases, model=model)
        self.sigdir = os.path.join(self.dir, "sig")
        os.makedirs(self.sigdir, exist_ok=True)
        self.reads.write_blow5(os.path.join(self.sigdir, "reads.blow5"))

    def pa(self, port, r):
        return port.raw_to_pa(self.reads.read(r), 8192.0, 10.0, 1437.976685)


@pytest.fixture(scope="session")
def small(host, model, tmp_path_factory):
    return Dataset(host, model, tmp_path_factory.mktemp("small"), [200000, 150000], 60)


@pytest.fixture(scope="session")
def mapper(small):
    from sigmap_b200.mapper import Mapper
    m = Mapper(0)
    m.set_index(small.pos, small.val)
    m.set_contigs(small.ref.lengths)
    yield m
    m.close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_chains(a, b):
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        for k in ("contig", "start", "end", "n_anchors", "mapq", "dir"):
            if int(x[k]) != int(y[k]):
                return False
        if bits(np.float32(x["score"])) != bits(np.float32(y["score"])):
            return False
        if len(x["anchors"]) != len(y["anchors"]):
            return False
        for (t1, q1, d1), (t2, q2, d2) in zip(x["anchors"], y["anchors"]):
            if int(t1) != int(t2) or int(q1) != int(q2) or bits(np.float32(d1)) != bits(np.float32(d2)):
                return False
    return True


def paf_cols(line):
    """PAF row without the wall-clock tag mt (compare everything else)."""
    return [c for c in line.rstrip("\n").split("\t") if not c.startswith("mt:f:")]


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class Golden:
    """tests/golden/: outputs of the unmodified reference on committed inputs
    (tests/make_golden.py)."""

    def __init__(self):
        import json
        self.z = np.load(os.path.join(GOLDEN_DIR, "stages.npz"))
        self.paf = json.load(open(os.path.join(GOLDEN_DIR, "paf.json")))

    def __getitem__(self, k):
        return self.z[k]

    def genome(self, host):
        lens = self.z["contig_len"]
        seq = self.z["contig_seq"].tobytes()
        o, seqs = 0, []
        for l in lens:
            seqs.append(seq[o:o + int(l)])
            o += int(l)
        return host.Reference([f"contig_{i}" for i in range(len(lens))], seqs)

    def reads(self, host):
        return host.ReadSet([str(n) for n in self.z["read_names"]], self.z["raw"], self.z["read_off"],
                            host.DIGITISATION, host.RANGE, host.OFFSET, self.z["truth"])

    def real_read(self, r):
        o = self.z["real_off"]
        return (self.z["real_raw"][int(o[r]):int(o[r + 1])], float(self.z["real_dig"][r]),
                float(self.z["real_offset"][r]), float(self.z["real_range"][r]))

    def chunk_features(self, ci):
        o = self.z["feat_off"]
        return self.z["feat"][int(o[ci]):int(o[ci + 1])]

    def hits(self, name, k):
        o = self.z[f"hits_{name}_off"]
        s = slice(int(o[k]), int(o[k + 1]))
        return self.z[f"hits_{name}_idx"][s], self.z[f"hits_{name}_d2"][s]

    def chain_states(self):
        """yield (read, chunk, [chain dicts]) in generation order"""
        ro = ao = 0
        for r, c, n_rec, n_anc in self.z["chain_key"]:
            rec = self.z["chain_rec"][ro:ro + n_rec]
            anc = self.z["chain_anc"][ao:ao + n_anc]
            ro += int(n_rec)
            ao += int(n_anc)
            chains, a0 = [], 0
            for row in rec:
                na = int(row[7])
                chains.append(dict(score=row[0:1].view(np.float32)[0], contig=int(row[1]),
                                   start=int(row[2]), end=int(row[3]), n_anchors=int(row[4]),
                                   mapq=int(row[5]), dir=int(row[6]),
                                   anchors=[(int(t), int(q), np.array([d], np.uint32).view(np.float32)[0])
                                            for t, q, d in anc[a0:a0 + na]]))
                a0 += na
            yield int(r), int(c), chains


@pytest.fixture(scope="session")
def golden():
    return Golden()
