"""Host-side helpers (no GPU): file formats and the synthetic-data simulator.

Thin numpy wrappers over the smbh_* functions of the C ABI; all the work is done in the
shared library (sigmap_b200/csrc/host_*.cc).
"""
import ctypes as C
import os

import numpy as np

from . import _ffi as F

MODEL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                          "r9.4_180mv_450bps_6mer_template_median68pA.model")

DIGITISATION, RANGE, OFFSET, SAMPLING_RATE = 8192.0, 1437.976685, 10.0, 4000.0


def _check(rc, what):
    if rc != 0:
        why = F.lib.smbh_last_error().decode()
        raise RuntimeError(f"{what} failed with code {rc}" + (f": {why}" if why else ""))


def load_pore_model(path=MODEL_PATH):
    """-> (level_mean[4096], level_stdv[4096]) fp32, indexed by the 2-bit 6-mer hash."""
    mean = np.zeros(4096, np.float32)
    stdv = np.zeros(4096, np.float32)
    _check(F.lib.smbh_pore_model_load(path.encode(), F.ptr(mean, F.f32p), F.ptr(stdv, F.f32p)),
           f"smbh_pore_model_load({path})")
    return mean, stdv


class Reference:
    """Contig names + sequences (bytes, NUL-terminated buffers kept alive for the C side)."""

    def __init__(self, names, seqs):
        self.names = list(names)
        self.seqs = [bytes(s) for s in seqs]
        self.lengths = np.array([len(s) for s in self.seqs], np.uint32)
        self._bufs = [C.create_string_buffer(s, len(s) + 1) for s in self.seqs]
        self.seq_ptrs = (C.c_char_p * len(self.seqs))(*[C.cast(b, C.c_char_p) for b in self._bufs])
        self.name_ptrs = (C.c_char_p * len(self.names))(*[n.encode() for n in self.names])

    @property
    def n(self):
        return len(self.names)

    def write_fasta(self, path):
        _check(F.lib.smbh_fasta_write(path.encode(), self.name_ptrs, self.seq_ptrs,
                                      F.ptr(self.lengths, F.u32p), self.n), "smbh_fasta_write")

    @staticmethod
    def read_fasta(path):
        fa = F.Fasta()
        _check(F.lib.smbh_fasta_load(path.encode(), C.byref(fa)), f"smbh_fasta_load({path})")
        try:
            names = [fa.names[i].decode() for i in range(fa.n)]
            seqs = [C.string_at(fa.seqs[i], fa.lengths[i]) for i in range(fa.n)]
        finally:
            F.lib.smbh_fasta_free(C.byref(fa))
        return Reference(names, seqs)


def sim_reference(seed, lengths, name_fmt="contig_{}"):
    lengths = np.asarray(lengths, np.uint32)
    bufs = [C.create_string_buffer(int(l) + 1) for l in lengths]
    ptrs = (C.c_char_p * len(bufs))(*[C.cast(b, C.c_char_p) for b in bufs])
    _check(F.lib.smbh_sim_reference(seed, F.ptr(lengths, F.u32p), len(bufs), ptrs),
           "smbh_sim_reference")
    return Reference([name_fmt.format(i) for i in range(len(bufs))], [b.raw[:-1] for b in bufs])


class ReadSet:
    """raw int16 samples of n reads, concatenated: read r = raw[read_off[r]:read_off[r+1]]."""

    def __init__(self, names, raw, read_off, digitisation, range_, offset, truth=None):
        self.names = list(names)
        self.raw = np.ascontiguousarray(raw, np.int16)
        self.read_off = np.ascontiguousarray(read_off, np.uint64)
        n = len(self.names)
        self.digitisation = np.broadcast_to(np.float32(digitisation), (n,)).astype(np.float32).copy()
        self.range = np.broadcast_to(np.float32(range_), (n,)).astype(np.float32).copy()
        self.offset = np.broadcast_to(np.float32(offset), (n,)).astype(np.float32).copy()
        self.truth = truth  # (n,4) uint32: contig, start, end, strand_plus

    @property
    def n(self):
        return len(self.names)

    def read(self, r):
        return self.raw[int(self.read_off[r]):int(self.read_off[r + 1])]

    def write_blow5(self, path):
        names = (C.c_char_p * self.n)(*[n.encode() for n in self.names])
        _check(F.lib.smbh_blow5_write(path.encode(), names, F.ptr(self.raw, F.i16p),
                                      F.ptr(self.read_off, F.u64p), self.n,
                                      float(self.digitisation[0]) if self.n else DIGITISATION,
                                      float(self.offset[0]) if self.n else OFFSET,
                                      float(self.range[0]) if self.n else RANGE, SAMPLING_RATE),
               "smbh_blow5_write")

    @staticmethod
    def read_blow5(path):
        r = F.Reads()
        _check(F.lib.smbh_blow5_read(path.encode(), C.byref(r)), f"smbh_blow5_read({path})")
        try:
            n = r.n
            off = np.ctypeslib.as_array(r.read_off, (n + 1,)).copy()
            raw = np.ctypeslib.as_array(r.raw, (max(int(off[-1]), 1),))[:int(off[-1])].copy()
            names = [r.names[i].decode() for i in range(n)]
            dig = np.ctypeslib.as_array(r.digitisation, (max(n, 1),))[:n].copy()
            rng = np.ctypeslib.as_array(r.range, (max(n, 1),))[:n].copy()
            ofs = np.ctypeslib.as_array(r.offset, (max(n, 1),))[:n].copy()
        finally:
            F.lib.smbh_reads_free(C.byref(r))
        return ReadSet(names, raw, off, dig, rng, ofs)


def sim_reads(seed, ref, n_reads, first_read=0, min_bases=2000, max_bases=9000, noise=1.0,
              model=None):
    """Simulate reads [first_read, first_read+n_reads) of the stream keyed by `seed`."""
    mean, stdv = model if model is not None else load_pore_model()
    off = np.zeros(n_reads + 1, np.uint64)
    truth = np.zeros((n_reads, 4), np.uint32)
    args = (seed, ref.seq_ptrs, F.ptr(ref.lengths, F.u32p), ref.n, F.ptr(mean, F.f32p),
            F.ptr(stdv, F.f32p), first_read, n_reads, min_bases, max_bases, noise)
    _check(F.lib.smbh_sim_reads(*args, F.ptr(off, F.u64p), None, F.ptr(truth, F.u32p)),
           "smbh_sim_reads(plan)")
    raw = np.zeros(int(off[-1]), np.int16)
    _check(F.lib.smbh_sim_reads(*args, F.ptr(off, F.u64p), F.ptr(raw, F.i16p), None),
           "smbh_sim_reads(fill)")
    names = ["read_%05d" % (first_read + i) for i in range(n_reads)]
    return ReadSet(names, raw, off, DIGITISATION, RANGE, OFFSET, truth)


def build_point_cloud(ref, level_mean):
    """Reference -> (pos uint64[n], val float32[n]) = the content of the reference's .pt."""
    pp, vp, n = F.u64p(), F.f32p(), C.c_size_t()
    _check(F.lib.smbh_build_point_cloud_alloc(ref.seq_ptrs, F.ptr(ref.lengths, F.u32p), ref.n,
                                              F.ptr(level_mean, F.f32p), C.byref(pp), C.byref(vp),
                                              C.byref(n)), "smbh_build_point_cloud_alloc")
    try:
        pos = np.ctypeslib.as_array(pp, (max(n.value, 1),))[:n.value].copy()
        val = np.ctypeslib.as_array(vp, (max(n.value, 1),))[:n.value].copy()
    finally:
        F.lib.smbh_free(pp)
        F.lib.smbh_free(vp)
    return pos, val


class CloudPart:
    """One rank's part of the point cloud (contig-sharded index): owns the C arrays."""

    def __init__(self, c):
        self.c = c

    def arrays(self):
        c = self.c
        nv, nr = c.n_values, c.n_runs
        take = lambda p, n, : np.ctypeslib.as_array(p, (max(n, 1),))[:n].copy()
        return dict(pos=take(c.pos, nv), val=take(c.val, nv), own=take(c.own, nv),
                    run_off=take(c.run_off, nr + 1), run_first=take(c.run_first, nr),
                    n_points_total=int(c.n_points_total))

    def close(self):
        if self.c is not None:
            F.lib.smbh_cloud_part_free(C.byref(self.c))
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_point_cloud_part(ref, level_mean, owner, rank):
    """The part of the reference's point cloud that shard `rank` needs (its own contigs' points
    and the five after each stretch): never the whole cloud in this process's memory."""
    owner = np.ascontiguousarray(owner, np.uint32)
    c = F.CloudPart()
    _check(F.lib.smbh_build_point_cloud_part(ref.seq_ptrs, F.ptr(ref.lengths, F.u32p), ref.n,
                                             F.ptr(level_mean, F.f32p), F.ptr(owner, F.u32p), rank,
                                             C.byref(c)), "smbh_build_point_cloud_part")
    return CloudPart(c)


def write_pt(prefix, pos, val, dim=6, max_leaf=20):
    pos = np.ascontiguousarray(pos, np.uint64)
    val = np.ascontiguousarray(val, np.float32)
    _check(F.lib.smbh_pt_write(prefix.encode(), F.ptr(pos, F.u64p), F.ptr(val, F.f32p), len(pos),
                               dim, max_leaf), "smbh_pt_write")


def write_si(prefix, val, dim=6, max_leaf=20):
    """<prefix>.si in nanoflann's layout, for the reference's own `sigmap -m` (we never read it)."""
    val = np.ascontiguousarray(val, np.float32)
    _check(F.lib.smbh_si_write(prefix.encode(), F.ptr(val, F.f32p), len(val), dim, max_leaf), "smbh_si_write")


def read_pt(prefix):
    pp, vp = F.u64p(), F.f32p()
    n, dim, ml = C.c_size_t(), C.c_int(), C.c_int()
    _check(F.lib.smbh_pt_read(prefix.encode(), C.byref(pp), C.byref(vp), C.byref(n), C.byref(dim),
                              C.byref(ml)), f"smbh_pt_read({prefix})")
    try:
        pos = np.ctypeslib.as_array(pp, (max(n.value, 1),))[:n.value].copy()
        val = np.ctypeslib.as_array(vp, (max(n.value, 1),))[:n.value].copy()
    finally:
        F.lib.smbh_free(pp)
        F.lib.smbh_free(vp)
    return pos, val, dim.value, ml.value


def format_paf(m, read_name, contig_name, contig_len, mt_ms=0.0):
    buf = C.create_string_buffer(2048)
    F.lib.smbh_format_paf(C.byref(m), read_name.encode(), contig_name.encode(), int(contig_len),
                          float(mt_ms), buf, 2048)
    return buf.value.decode()
