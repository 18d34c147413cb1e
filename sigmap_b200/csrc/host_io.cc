// host_io.cc -- host-side file formats on either side of the mapping hot path:
// pore-model TSV, FASTA, the reference's `.pt` point-cloud file, BLOW5 and the PAF row.
// No GPU code here; these are the smbh_* helpers of include/sigmap_b200.h.
//
// Formats follow the reference's readers/writers so files are interchangeable:
//   pore model  pore_model.cc:11-47      .pt   spatial_index.cc:105-147
//   FASTA       sequence_batch.cc (kseq) BLOW5 slow5lib 0.2.0 (extern/slow5lib/src/slow5.c)
//   PAF         output_tools.h:200-210,336-354 + sigmap.cc:731-745
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sigmap_b200.h"
#include "sb_host.h"

namespace sb {

int base_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

}  // namespace sb

extern "C" {

void smbh_free(void *p) { free(p); }

// ------------------------------------------------------------------ pore model
int smbh_pore_model_load(const char *path, float *level_mean, float *level_stdv) {
  FILE *f = fopen(path, "r");
  if (!f) return SMB_ERR_IO;
  for (int i = 0; i < 4096; ++i) {
    level_mean[i] = 0;
    if (level_stdv) level_stdv[i] = 0;
  }
  char line[1024];
  int n = 0;
  while (fgets(line, sizeof line, f)) {
    if (line[0] == '#' || strncmp(line, "kmer", 4) == 0) continue;  // pore_model.cc:19-23
    char kmer[64];
    double mean, stdv;
    if (sscanf(line, "%63s %lf %lf", kmer, &mean, &stdv) != 3) continue;
    if (strlen(kmer) != 6) {
      fclose(f);
      return SMB_ERR_IO;  // only the 6-mer models are on the supported path
    }
    uint32_t h = 0;
    for (int i = 0; i < 6; ++i) {
      int c = sb::base_code(kmer[i]);
      h = (h << 2) | (uint32_t)(c < 0 ? 0 : c);
    }
    level_mean[h] = (float)mean;
    if (level_stdv) level_stdv[h] = (float)stdv;
    ++n;
  }
  fclose(f);
  return n == 4096 ? SMB_OK : SMB_ERR_IO;
}

// ------------------------------------------------------------------ FASTA
int smbh_fasta_load(const char *path, smbh_fasta *out) {
  memset(out, 0, sizeof *out);
  gzFile f = gzopen(path, "r");
  if (!f) return SMB_ERR_IO;
  std::vector<std::string> names, seqs;
  std::vector<char> buf(4 << 20);
  // Line-wise over whole buffer segments (memchr + append) rather than byte by byte.  '\n' and
  // '\r' both end a line, a '>' at a line start opens a record whose name ends at the first
  // whitespace (kseq), every byte <= ' ' inside sequence lines is dropped, text before the first
  // header is ignored.
  std::string header;
  bool at_line_start = true, in_header = false, in_seq = false;
  auto finish_header = [&]() {
    size_t e = 0;
    while (e < header.size() && (unsigned char)header[e] > ' ') ++e;
    names.push_back(header.substr(0, e));
    seqs.emplace_back();
    in_seq = true;
    in_header = false;
  };
  auto append_bases = [&](const char *p, size_t n) {
    std::string &s = seqs.back();
    bool clean = true;
    for (size_t k = 0; k < n; ++k) clean &= (unsigned char)p[k] > ' ';
    if (clean) {
      s.append(p, n);
    } else {
      for (size_t k = 0; k < n; ++k)
        if ((unsigned char)p[k] > ' ') s.push_back(p[k]);
    }
  };
  int got;
  while ((got = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
    const char *b = buf.data();
    size_t i = 0;
    const size_t n = (size_t)got;
    while (i < n) {
      if (at_line_start) {
        at_line_start = false;
        if (b[i] == '\n' || b[i] == '\r') {  // empty line
          at_line_start = true;
          ++i;
          continue;
        }
        if (b[i] == '>') {
          in_header = true;
          header.clear();
          ++i;
          continue;
        }
      }
      const char *nl = (const char *)memchr(b + i, '\n', n - i);
      size_t e = nl ? (size_t)(nl - b) : n;
      const char *cr = (const char *)memchr(b + i, '\r', e - i);
      if (cr) e = (size_t)(cr - b);
      if (in_header) header.append(b + i, e - i);
      else if (in_seq) append_bases(b + i, e - i);
      if (e < n) {  // a line terminator at e
        if (in_header) finish_header();
        at_line_start = true;
        i = e + 1;
      } else {
        i = e;
      }
    }
  }
  if (in_header) finish_header();
  gzclose(f);
  // sequence_batch.cc:22-25: zero-length records are skipped
  std::vector<size_t> keep;
  for (size_t i = 0; i < seqs.size(); ++i)
    if (!seqs[i].empty()) keep.push_back(i);
  out->n = (uint32_t)keep.size();
  out->names = (char **)calloc(keep.size() + 1, sizeof(char *));
  out->seqs = (char **)calloc(keep.size() + 1, sizeof(char *));
  out->lengths = (uint32_t *)calloc(keep.size() + 1, sizeof(uint32_t));
  for (size_t k = 0; k < keep.size(); ++k) {
    size_t i = keep[k];
    out->names[k] = strdup(names[i].c_str());
    out->seqs[k] = (char *)malloc(seqs[i].size() + 1);
    memcpy(out->seqs[k], seqs[i].c_str(), seqs[i].size() + 1);
    out->lengths[k] = (uint32_t)seqs[i].size();
  }
  return SMB_OK;
}

void smbh_fasta_free(smbh_fasta *f) {
  if (!f) return;
  for (uint32_t i = 0; i < f->n; ++i) {
    free(f->names[i]);
    free(f->seqs[i]);
  }
  free(f->names);
  free(f->seqs);
  free(f->lengths);
  memset(f, 0, sizeof *f);
}

int smbh_fasta_write(const char *path, const char *const *names, const char *const *seqs,
                     const uint32_t *lengths, uint32_t n) {
  FILE *f = fopen(path, "w");
  if (!f) return SMB_ERR_IO;
  for (uint32_t i = 0; i < n; ++i) {
    fprintf(f, ">%s\n", names[i]);
    for (uint32_t p = 0; p < lengths[i]; p += 60) {
      uint32_t w = lengths[i] - p < 60 ? lengths[i] - p : 60;
      fwrite(seqs[i] + p, 1, w, f);
      fputc('\n', f);
    }
  }
  fclose(f);
  return SMB_OK;
}

// ------------------------------------------------------------------ .pt
// spatial_index.cc:105-123: int dim, int max_leaf, size_t n, n x Point{u64 pos; f32 value; 4 B pad}
int smbh_pt_write(const char *prefix, const uint64_t *pos, const float *val, size_t n, int dim,
                  int max_leaf) {
  std::string p = std::string(prefix) + ".pt";
  FILE *f = fopen(p.c_str(), "wb");
  if (!f) return SMB_ERR_IO;
  uint64_t n64 = n;
  fwrite(&dim, sizeof(int), 1, f);
  fwrite(&max_leaf, sizeof(int), 1, f);
  fwrite(&n64, sizeof(uint64_t), 1, f);
  struct Rec {
    uint64_t pos;
    float val;
    uint32_t pad;
  };
  std::vector<Rec> buf(1 << 16);
  for (size_t i = 0; i < n;) {
    size_t m = n - i < buf.size() ? n - i : buf.size();
    for (size_t k = 0; k < m; ++k) buf[k] = Rec{pos[i + k], val[i + k], 0};
    if (fwrite(buf.data(), sizeof(Rec), m, f) != m) {
      fclose(f);
      return SMB_ERR_IO;
    }
    i += m;
  }
  fclose(f);
  return SMB_OK;
}

int smbh_pt_read(const char *prefix, uint64_t **pos, float **val, size_t *n, int *dim,
                 int *max_leaf) {
  // mapped file, records de-interleaved on all host cores (a 3.1 Gbp index is a 99 GB file)
  std::string p = std::string(prefix) + ".pt";
  const int fd = open(p.c_str(), O_RDONLY);
  if (fd < 0) return SMB_ERR_IO;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 16) {
    close(fd);
    return SMB_ERR_IO;
  }
  const size_t fsize = (size_t)st.st_size;
  void *map = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return SMB_ERR_IO;
  const unsigned char *base = static_cast<const unsigned char *>(map);
  int d = 0, ml = 0;
  uint64_t n64 = 0;
  memcpy(&d, base, 4);
  memcpy(&ml, base + 4, 4);
  memcpy(&n64, base + 8, 8);
  struct Rec {
    uint64_t pos;
    float val;
    uint32_t pad;
  };
  if (n64 > (fsize - 16) / sizeof(Rec)) {  // truncated file: fail loudly
    munmap(map, fsize);
    return SMB_ERR_IO;
  }
  uint64_t *ps = (uint64_t *)malloc((n64 ? n64 : 1) * sizeof(uint64_t));
  float *vs = (float *)malloc((n64 ? n64 : 1) * sizeof(float));
  if (!ps || !vs) {
    free(ps);
    free(vs);
    munmap(map, fsize);
    return SMB_ERR_IO;
  }
  const Rec *rec = reinterpret_cast<const Rec *>(base + 16);  // 16-byte header: records stay aligned
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n64; ++i) {
    ps[i] = rec[i].pos;
    vs[i] = rec[i].val;
  }
  munmap(map, fsize);
  *pos = ps;
  *val = vs;
  *n = (size_t)n64;
  if (dim) *dim = d;
  if (max_leaf) *max_leaf = ml;
  return SMB_OK;
}

// ------------------------------------------------------------------ BLOW5
// slow5lib 0.2.0 binary layout (extern/slow5lib/src/slow5.c:520-600, :697-850):
//   "BLOW5\1" | u8 major,minor,patch | u8 compression (0 none, 1 zlib) | u32 n_read_groups
//   | zero pad to byte 64 | u32 ascii_header_len | ascii header
//   records: u64 rec_bytes | [u16 id_len | id | u32 read_group | f64 digitisation | f64 offset
//            | f64 range | f64 sampling_rate | u64 n | int16[n] | aux...]   (zlib: the
//            bracketed part is one deflate stream)
//   | "5WOLB"
int smbh_blow5_write(const char *path, const char *const *names, const int16_t *raw,
                     const uint64_t *read_off, size_t n, double digitisation, double offset,
                     double range, double sampling_rate) {
  FILE *f = fopen(path, "wb");
  if (!f) return SMB_ERR_IO;
  unsigned char head[64];
  memset(head, 0, sizeof head);
  memcpy(head, "BLOW5\1", 6);
  head[6] = 0;
  head[7] = 1;
  head[8] = 0;   // file version 0.1.0
  head[9] = 0;   // no compression
  uint32_t nrg = 1;
  memcpy(head + 10, &nrg, 4);
  fwrite(head, 1, 64, f);
  char hdr[512];
  int hl = snprintf(hdr, sizeof hdr,
                    "@asic_id\t0\n@sample_frequency\t%d\n"
                    "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
                    "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\t"
                    "len_raw_signal\traw_signal\n",
                    (int)sampling_rate);
  uint32_t hl32 = (uint32_t)hl;
  fwrite(&hl32, 4, 1, f);
  fwrite(hdr, 1, hl, f);
  for (size_t r = 0; r < n; ++r) {
    uint16_t idl = (uint16_t)strlen(names[r]);
    uint64_t ns = read_off[r + 1] - read_off[r];
    uint64_t rec = 2 + idl + 4 + 8 * 4 + 8 + 2 * ns;
    uint32_t rg = 0;
    fwrite(&rec, 8, 1, f);
    fwrite(&idl, 2, 1, f);
    fwrite(names[r], 1, idl, f);
    fwrite(&rg, 4, 1, f);
    fwrite(&digitisation, 8, 1, f);
    fwrite(&offset, 8, 1, f);
    fwrite(&range, 8, 1, f);
    fwrite(&sampling_rate, 8, 1, f);
    fwrite(&ns, 8, 1, f);
    if (fwrite(raw + read_off[r], 2, ns, f) != ns) {
      fclose(f);
      return SMB_ERR_IO;
    }
  }
  fwrite("5WOLB", 1, 5, f);
  fclose(f);
  return SMB_OK;
}

static bool inflate_all(const unsigned char *src, size_t n, std::vector<unsigned char> &out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit(&zs) != Z_OK) return false;
  zs.next_in = const_cast<unsigned char *>(src);
  zs.avail_in = (uInt)n;
  out.resize(n * 4 + 1024);
  size_t done = 0;
  int ret;
  do {
    if (done == out.size()) out.resize(out.size() * 2);
    zs.next_out = out.data() + done;
    zs.avail_out = (uInt)(out.size() - done);
    ret = inflate(&zs, Z_NO_FLUSH);
    done = out.size() - zs.avail_out;
  } while (ret == Z_OK);
  inflateEnd(&zs);
  out.resize(done);
  return ret == Z_STREAM_END;
}

// The file is mapped, the record boundaries are found in one sequential walk over the 8-byte
// length prefixes, and everything per record -- inflating zlib records, parsing the header,
// copying the samples to their final place -- runs on all host cores.  The raw samples are
// copied exactly once, into an array sized from the record headers.
static thread_local std::string g_host_error;
static int host_fail(int code, const std::string &msg) {
  g_host_error = msg;
  return code;
}
const char *smbh_last_error(void) { return g_host_error.c_str(); }

int smbh_blow5_read(const char *path, smbh_reads *out) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return host_fail(SMB_ERR_IO, std::string("cannot open ") + path);
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 68) {
    close(fd);
    return SMB_ERR_IO;
  }
  const size_t fsize = (size_t)st.st_size;
  void *map = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return SMB_ERR_IO;
  const unsigned char *base = static_cast<const unsigned char *>(map);
  struct Unmap {
    void *p;
    size_t n;
    ~Unmap() { munmap(p, n); }
  } unmap{map, fsize};
  if (memcmp(base, "BLOW5\1", 6) != 0) return host_fail(SMB_ERR_IO, std::string(path) + ": not a BLOW5 file");
  // version triple, then: 0.1.0 (the reference's bundled slow5lib) {record compression u8,
  // read groups u32}; from 0.2.0 on {record compression u8, SIGNAL compression u8, read groups u32}
  const int vmaj = base[6], vmin = base[7], vpat = base[8];
  const std::string ver = std::to_string(vmaj) + "." + std::to_string(vmin) + "." + std::to_string(vpat);
  if (vmaj != 0 || (vmin != 1 && vmin != 2)) return host_fail(SMB_ERR_IO, std::string(path) + ": unsupported BLOW5 version " + ver);
  const int method = base[9];
  if (vmin >= 2 && base[10] != 0)
    return host_fail(SMB_ERR_IO, std::string(path) + ": unsupported BLOW5 signal compression (method " +
                                     std::to_string((int)base[10]) + ", version " + ver + "); only uncompressed signals are read");
  uint32_t hl;
  memcpy(&hl, base + 64, 4);
  if (method > 1) return host_fail(SMB_ERR_IO, std::string(path) + ": unsupported BLOW5 record compression (only none and zlib)");
  if ((size_t)68 + hl > fsize) return host_fail(SMB_ERR_IO, std::string(path) + ": truncated BLOW5 header");

  // ---- record boundaries
  std::vector<size_t> rec_at;
  std::vector<uint64_t> rec_len;
  size_t at = (size_t)68 + hl;
  for (;;) {
    if (at + 5 <= fsize && memcmp(base + at, "5WOLB", 5) == 0 && at + 5 == fsize) break;  // EOF marker
    if (at + 8 > fsize) return SMB_ERR_IO;  // truncated (no EOF marker)
    uint64_t rl;
    memcpy(&rl, base + at, 8);
    if (rl > fsize - at - 8) return SMB_ERR_IO;
    rec_at.push_back(at + 8);
    rec_len.push_back(rl);
    at += 8 + rl;
  }
  const size_t nrec = rec_at.size();

  // ---- per record: plain bytes (inflated when needed) and the header fields
  std::vector<std::vector<unsigned char>> plain(method == 1 ? nrec : 0);
  std::vector<const unsigned char *> body(nrec, nullptr);  // first sample
  std::vector<const unsigned char *> name_at(nrec, nullptr);
  std::vector<uint16_t> name_len(nrec, 0);
  std::vector<uint64_t> ns(nrec, 0);
  std::vector<float> dig(nrec), rng(nrec), off(nrec);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad)
  for (int64_t i = 0; i < (int64_t)nrec; ++i) {
    const unsigned char *p = base + rec_at[i];
    size_t pn = rec_len[i];
    if (method == 1) {
      if (!inflate_all(p, pn, plain[i])) {
        bad |= 1;
        continue;
      }
      p = plain[i].data();
      pn = plain[i].size();
    }
    if (pn < 2) { bad |= 1; continue; }
    uint16_t idl;
    memcpy(&idl, p, 2);
    if (pn < (size_t)2 + idl + 4 + 32 + 8) { bad |= 1; continue; }
    const unsigned char *q = p + 2 + idl + 4;
    double d4[4];
    memcpy(d4, q, 32);
    uint64_t n;
    memcpy(&n, q + 32, 8);
    if (n > (pn - ((size_t)2 + idl + 4 + 32 + 8)) / 2) { bad |= 1; continue; }
    name_at[i] = p + 2;
    name_len[i] = idl;
    // signal_batch.cc:187-191 narrows the doubles to float
    dig[i] = (float)d4[0];
    off[i] = (float)d4[1];
    rng[i] = (float)d4[2];
    ns[i] = n;
    body[i] = q + 40;
  }
  if (bad) return SMB_ERR_IO;

  // ---- append to *out
  const size_t n0 = out->n, n1 = n0 + nrec;
  const uint64_t s0 = n0 ? out->read_off[n0] : 0;
  uint64_t added = 0;
  for (size_t i = 0; i < nrec; ++i) added += ns[i];
  out->names = (char **)realloc(out->names, (n1 + 1) * sizeof(char *));
  out->read_off = (uint64_t *)realloc(out->read_off, (n1 + 1) * sizeof(uint64_t));
  out->digitisation = (float *)realloc(out->digitisation, (n1 + 1) * sizeof(float));
  out->range = (float *)realloc(out->range, (n1 + 1) * sizeof(float));
  out->offset = (float *)realloc(out->offset, (n1 + 1) * sizeof(float));
  out->raw = (int16_t *)realloc(out->raw, (s0 + added + 1) * sizeof(int16_t));
  if (!out->names || !out->read_off || !out->digitisation || !out->range || !out->offset || !out->raw)
    return SMB_ERR_IO;
  uint64_t run = s0;
  for (size_t i = 0; i < nrec; ++i) {
    out->read_off[n0 + i] = run;
    run += ns[i];
  }
  out->read_off[n1] = run;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < (int64_t)nrec; ++i) {
    memcpy(out->raw + out->read_off[n0 + i], body[i], 2 * ns[i]);
    out->names[n0 + i] = strndup((const char *)name_at[i], name_len[i]);
    out->digitisation[n0 + i] = dig[i];
    out->range[n0 + i] = rng[i];
    out->offset[n0 + i] = off[i];
  }
  out->n = n1;
  return SMB_OK;
}

void smbh_reads_free(smbh_reads *r) {
  if (!r) return;
  for (size_t i = 0; i < r->n; ++i) free(r->names[i]);
  free(r->names);
  free(r->read_off);
  free(r->raw);
  free(r->digitisation);
  free(r->range);
  free(r->offset);
  memset(r, 0, sizeof *r);
}

// ------------------------------------------------------------------ PAF
// std::to_string(float/double) is "%f"; mapped rows: output_tools.h:336-354, unmapped rows:
// :200-210 (nine '*' columns, mapq 61); tags: sigmap.cc:731-745 / :826-858.
int smbh_format_paf(const smb_mapping *m, const char *read_name, const char *contig_name,
                    uint32_t contig_len, double mt_ms, char *buf, size_t cap) {
  char tags[640];
  int k = snprintf(tags, sizeof tags, "mt:f:%f\tci:i:%u\tsl:i:%u", mt_ms, m->chunks, m->read_len);
  if (m->n_chains >= 1)
    snprintf(tags + k, sizeof tags - k,
             "\tcm:i:%u\tnc:i:%u\ts1:f:%f\ts2:f:%f\tsm:f:%f\tad:f:%f\tat:f:%f\taq:f:%f", m->cm,
             m->n_chains, (double)m->s1, (double)m->s2, (double)m->sm, (double)m->ad,
             (double)m->at, (double)m->aq);
  if (m->mapped && m->mapq <= 60)
    return snprintf(buf, cap, "%s\t%u\t%u\t%u\t%s\t%s\t%u\t%u\t%u\t%u\t%u\t%u\t%s\n", read_name,
                    m->read_len, m->q_start, m->q_end, m->strand_plus ? "+" : "-", contig_name,
                    contig_len, m->t_start, m->t_start + m->frag_len, m->read_len, m->frag_len,
                    m->mapq, tags);
  return snprintf(buf, cap, "%s\t%u\t*\t*\t*\t*\t*\t*\t*\t*\t*\t%u\t%s\n", read_name, m->read_len,
                  61u, tags);
}

}  // extern "C"
