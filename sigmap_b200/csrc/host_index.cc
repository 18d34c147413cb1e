// host_index.cc -- reference sequence -> expected-current point cloud, i.e. the data half
// of `sigmap -i` (Sigmap::ConstructIndex, sigmap.cc:999-1046).  The output is the exact
// content of the reference's `.pt` file, which is what the device index is built from.
//
// Stages and the reference lines they restate:
//   6-mer level lookup incl. its read-one-base-too-far quirk   pore_model.cc:57-80 (Q1)
//   per-strand z-normalisation (double accumulators, n-1)       sigmap.cc:1131-1155
//   high-frequency 11-mer masking (canonical k-mer counts)      sigmap.cc:19-185, :1014
//   point selection: skip masked, skip |dz| <= 0.01 vs last     spatial_index.cc:33-57
//   order: all + strands by contig, then all - strands          spatial_index.cc:82-93
//   packed position ((contig<<32 | p) << 1) | strand            spatial_index.cc:47-51
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sigmap_b200.h"
#include "sb_host.h"

// Everything position-parallel runs on all host cores (OpenMP): the reverse complement, the
// expected levels, the 11-mer census and masking, the final division of the z-scores.  What the
// reference's arithmetic makes order-dependent stays sequential and identical: the double
// accumulators of the z-normalisation and the `last kept value` chain of the point selection.
// A 3.1 Gbp reference is ~6.2 G positions; at one core this stage took ~18 minutes.
namespace {

constexpr int kK = 6;                    // pore-model k
constexpr int kMaskK = SMB_DIM + kK - 1; // 11-mers are masked (sigmap.cc:1014)
constexpr float kMaskFreq = 0.0002f;
constexpr double kMinDelta = 0.01;       // spatial_index.cc:46
constexpr uint32_t kChunk = 1u << 20;    // positions per parallel work item

struct Codes {  // A=0 C=1 G=2 T=3 (either case), anything else -1: sb::base_code as a table
  signed char t[256];
  Codes() {
    for (int c = 0; c < 256; ++c) t[c] = (signed char)sb::base_code((char)c);
  }
  int operator()(char c) const { return t[(unsigned char)c]; }
};
const Codes kCode;

struct StrandView {
  std::string owned;  // reverse complement when needed
  const char *s;
  uint32_t len;
};

void make_strand(StrandView &v, const char *seq, uint32_t len, int strand) {
  v.len = len;
  if (strand == 0) {
    v.s = seq;
    return;
  }
  v.owned.resize(len);
  char *out = &v.owned[0];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)len; ++i) {  // sequence_batch.h:66-77
    const int c = kCode(seq[len - 1 - i]);
    out[i] = c < 0 ? 'N' : "ACGT"[3 ^ c];
  }
  v.s = v.owned.c_str();  // NUL-terminated like std::string::data() in the reference
}

// Expected current per position.  Position 0 hashes bases 0..5; position p >= 1 shifts in
// base p+6 (not p+5), so the last position shifts in the terminating NUL, read as 'A'.  From
// p = 6 on the hash therefore holds bases p+1..p+6, which lets chunks start anywhere.
void expected_levels(const StrandView &v, const float *level_mean, std::vector<float> &out) {
  const uint32_t L = v.len - kK + 1, mask = (1u << (2 * kK)) - 1;
  out.resize(L);
  auto code0 = [&](uint32_t i) {  // ambiguous bases (and the NUL) hash as 'A'
    const int c = kCode(v.s[i]);
    return (uint32_t)(c < 0 ? 0 : c);
  };
  uint32_t h = 0;
  for (uint32_t i = 0; i < (uint32_t)kK; ++i) h = ((h << 2) | (i < L ? code0(i) : 0u)) & mask;
  out[0] = level_mean[h];
  const uint32_t head = L < 7u ? L : 7u;
  for (uint32_t p = 1; p < head; ++p) {
    h = ((h << 2) | code0(p + kK)) & mask;
    out[p] = level_mean[h];
  }
  float *o = out.data();
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t a = 7; a < (int64_t)L; a += kChunk) {
    const uint32_t b = (uint32_t)std::min<int64_t>(L, a + kChunk);
    uint32_t g = 0;
    for (uint32_t i = (uint32_t)a + 1; i < (uint32_t)a + kK; ++i) g = (g << 2) | code0(i);  // bases a+1..a+5
    for (uint32_t q = (uint32_t)a; q < b; ++q) {  // g holds bases q+1..q+5: shift in base q+6
      g = ((g << 2) | code0(q + kK)) & mask;      // now bases q+1..q+6
      o[q] = level_mean[g];
    }
  }
}

void znormalise(std::vector<float> &x) {
  const size_t n = x.size();
  double mean = 0;
  for (size_t i = 0; i < n; ++i) mean += x[i];
  mean /= n;
  double ss = 0;
  for (size_t i = 0; i < n; ++i) ss += (x[i] - mean) * (x[i] - mean);
  const double sd = std::sqrt(ss / (n - 1));
  float *p = x.data();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) p[i] = (float)((p[i] - mean) / sd);
}

// Walks the k-mers STARTING in [a, b) of a strand; fn(pos_of_kmer_start, canonical_kmer) for
// complete k-mers and fn_amb(pos) where the window *ending* at an ambiguous base starts.
// Equivalent to one walk over the whole strand restricted to those starts: the run counter and
// both hashes only depend on the last 11 bases.
template <class F, class G>
void walk_kmers(const StrandView &v, uint32_t a, uint32_t b, F fn, G fn_amb) {
  const uint64_t m = (1ull << (2 * kMaskK)) - 1, shift = 2ull * (kMaskK - 1);
  uint64_t fw = 0, rv = 0;
  int run = 0;
  const uint64_t end = std::min<uint64_t>(v.len, (uint64_t)b + kMaskK - 1);
  for (uint64_t p = a; p < end; ++p) {
    const int c = kCode(v.s[p]);
    if (c >= 0) {
      fw = ((fw << 2) | (uint64_t)c) & m;
      rv = (rv >> 2) | ((uint64_t)(3 ^ c) << shift);
      if (++run >= kMaskK) fn((uint32_t)(p + 1 - kMaskK), fw < rv ? fw : rv);
    } else {
      run = 0;
      fw = rv = 0;
      if (p >= (uint64_t)a + kMaskK - 1) fn_amb((uint32_t)(p + 1 - kMaskK));
    }
  }
}

// Where the points go.  Whole cloud: pos/val (either may be null: count only).  One rank's part of a
// contig-sharded index (part != nullptr): only the points of its own contigs plus, after each
// stretch of them, the kDim-1 points that follow in the whole cloud (the windows at the end of
// the stretch straddle into them, Q2) -- so a rank never holds more than its share of a genome-
// scale cloud.
struct PartSink {
  const uint32_t *owner;
  uint32_t rank;
  std::vector<uint64_t> pos, run_off, run_first;
  std::vector<float> val;
  std::vector<unsigned char> own;
  int trailing = 0;      // points after the last owned one still to be taken
  bool open = false;     // a run is being written
  void point(uint32_t contig, uint64_t P, float z, uint64_t index) {
    const bool mine = owner[contig] == rank;
    if (!mine && trailing == 0) {
      open = false;
      return;
    }
    if (!open) {
      run_off.push_back(pos.size());
      run_first.push_back(index);
      open = true;
    }
    pos.push_back(P);
    val.push_back(z);
    own.push_back(mine ? 1 : 0);
    trailing = mine ? SMB_DIM - 1 : trailing - 1;
    if (!mine && trailing == 0) open = false;
  }
};

size_t build_cloud(const char *const *seqs, const uint32_t *lengths, uint32_t n, const float *level_mean,
                   uint64_t *pos, float *val, PartSink *sink = nullptr) {
  // k-mer census over the + strands (canonical, so the - strand adds nothing new)
  std::vector<uint32_t> hist((size_t)1 << (2 * kMaskK), 0);
  uint32_t *H = hist.data();
  uint64_t total = 0;
  for (uint32_t s = 0; s < n; ++s) {
    StrandView v;
    make_strand(v, seqs[s], lengths[s], 0);
    uint64_t part = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : part)
    for (int64_t a = 0; a < (int64_t)v.len; a += kChunk) {
      uint64_t mine = 0;
      walk_kmers(v, (uint32_t)a, (uint32_t)std::min<int64_t>(v.len, a + kChunk),
                 [&](uint32_t, uint64_t k) {
#pragma omp atomic
                   ++H[k];
                   ++mine;
                 },
                 [](uint32_t) {});
      part += mine;
    }
    total += part;
  }
  size_t count = 0;
  bool any = false;
  float last = 0;
  std::vector<float> z;
  std::vector<unsigned char> masked;
  StrandView v;
  for (int strand = 0; strand < 2; ++strand) {
    for (uint32_t s = 0; s < n; ++s) {
      if (lengths[s] < (uint32_t)kMaskK) continue;
      make_strand(v, seqs[s], lengths[s], strand);
      expected_levels(v, level_mean, z);
      znormalise(z);
      const uint32_t n_starts = v.len - kMaskK + 1;
      masked.assign(n_starts, 0);
      unsigned char *M = masked.data();
#pragma omp parallel for schedule(dynamic, 1)
      for (int64_t a = 0; a < (int64_t)n_starts; a += kChunk)
        walk_kmers(
            v, (uint32_t)a, (uint32_t)std::min<int64_t>(n_starts, a + kChunk),
            [&](uint32_t p, uint64_t k) { M[p] = ((float)H[k] / (float)total) > kMaskFreq; },
            [&](uint32_t p) { M[p] = 1; });
      const uint32_t n_windows = (uint32_t)z.size() - SMB_DIM + 1;
      for (uint32_t p = 0; p < n_windows; ++p) {
        if (masked[p]) continue;
        if (p == 0 || !any || std::fabs((double)(z[p] - last)) > kMinDelta) {
          const uint64_t P = ((((uint64_t)s << 32) | p) << 1) | (uint64_t)strand;
          if (sink) {
            sink->point(s, P, z[p], count);
          } else if (pos) {
            pos[count] = P;
            val[count] = z[p];
          }
          last = z[p];
          any = true;
          ++count;
        }
      }
    }
  }
  return count;
}

}  // namespace

extern "C" size_t smbh_build_point_cloud(const char *const *seqs, const uint32_t *lengths,
                                         uint32_t n, const float *level_mean, uint64_t *pos,
                                         float *val) {
  return build_cloud(seqs, lengths, n, level_mean, pos, val);
}

// One pass instead of count + fill: the arrays are allocated at the upper bound (one point per
// window; untouched pages cost nothing) and trimmed.  Release with smbh_free.
extern "C" int smbh_build_point_cloud_alloc(const char *const *seqs, const uint32_t *lengths, uint32_t n,
                                            const float *level_mean, uint64_t **pos, float **val,
                                            size_t *count) {
  size_t upper = 1;
  for (uint32_t s = 0; s < n; ++s)
    if (lengths[s] >= (uint32_t)kMaskK) upper += 2 * (size_t)(lengths[s] - kMaskK + 1);
  uint64_t *P = (uint64_t *)malloc(upper * sizeof(uint64_t));
  float *V = (float *)malloc(upper * sizeof(float));
  if (!P || !V) {
    free(P);
    free(V);
    return SMB_ERR_IO;
  }
  const size_t c = build_cloud(seqs, lengths, n, level_mean, P, V);
  uint64_t *P2 = (uint64_t *)realloc(P, std::max<size_t>(c, 1) * sizeof(uint64_t));
  float *V2 = (float *)realloc(V, std::max<size_t>(c, 1) * sizeof(float));
  *pos = P2 ? P2 : P;
  *val = V2 ? V2 : V;
  *count = c;
  return SMB_OK;
}

// One rank's part of the cloud (contig-sharded index, SURVEY.md 8e mode 2).  The census and the
// `last kept value` chain still run over the whole genome -- they decide which points exist --
// but only the rank's own points (and the five after each stretch) are kept in memory.
extern "C" int smbh_build_point_cloud_part(const char *const *seqs, const uint32_t *lengths, uint32_t n,
                                           const float *level_mean, const uint32_t *owner, uint32_t rank,
                                           smbh_cloud_part *out) {
  memset(out, 0, sizeof *out);
  if (!owner) return SMB_ERR_ARG;
  PartSink sink;
  sink.owner = owner;
  sink.rank = rank;
  const size_t total = build_cloud(seqs, lengths, n, level_mean, nullptr, nullptr, &sink);
  sink.run_off.push_back(sink.pos.size());
  out->n_points_total = total;
  out->n_values = sink.pos.size();
  out->n_runs = sink.run_first.size();
  auto dup = [](const void *src, size_t bytes) -> void * {
    void *p = malloc(bytes ? bytes : 1);
    if (p && bytes) memcpy(p, src, bytes);
    return p;
  };
  out->pos = (uint64_t *)dup(sink.pos.data(), sink.pos.size() * sizeof(uint64_t));
  out->val = (float *)dup(sink.val.data(), sink.val.size() * sizeof(float));
  out->own = (uint8_t *)dup(sink.own.data(), sink.own.size());
  out->run_off = (uint64_t *)dup(sink.run_off.data(), sink.run_off.size() * sizeof(uint64_t));
  out->run_first = (uint64_t *)dup(sink.run_first.data(), sink.run_first.size() * sizeof(uint64_t));
  if (!out->pos || !out->val || !out->own || !out->run_off || !out->run_first) {
    smbh_cloud_part_free(out);
    return SMB_ERR_IO;
  }
  return SMB_OK;
}

extern "C" void smbh_cloud_part_free(smbh_cloud_part *p) {
  if (!p) return;
  free(p->pos);
  free(p->val);
  free(p->own);
  free(p->run_off);
  free(p->run_first);
  memset(p, 0, sizeof *p);
}

// ---------------------------------------------------------------------------------------------
// <prefix>.si: a KD-tree over the window points in the on-disk layout of nanoflann 1.3.2's
// saveIndex_ (nanoflann.hpp:1051-1058, written by SpatialIndex::Save, spatial_index.cc:105-130), so
// that an index built by this program can be loaded by the reference's `sigmap -m`:
//   size_t m_size (= points - dim + 1), int dim, vector<{float low, high}> root_bbox,
//   size_t leaf_max_size, vector<size_t> vind, then the nodes in pre-order, 32 bytes each:
//   union {leaf: size_t left, right | inner: int divfeat; float divlow, divhigh}, two child pointers
//   (only tested for NULL by load_tree, nanoflann.hpp:1035-1044).
// The tree is this program's own (balanced median splits on the widest dimension; nanoflann splits
// at the middle value): searchLevel (nanoflann.hpp:1347-1410) only needs every point of child1 to
// be <= divlow and every point of child2 >= divhigh in dimension divfeat, and the exact radius
// search returns the same set from any valid tree.  (The order of the hits, which the 5 000-hit cap
// of spatial_index.cc:371-372 depends on, is the tree's -- as it is for any other builder.)
namespace {

struct SiNode {  // sizeof == 32, as nanoflann's Node on LP64
  union {
    struct { uint64_t left, right; } lr;
    struct { int32_t divfeat; float divlow, divhigh; } sub;
  } u;
  uint64_t child1, child2;
};
static_assert(sizeof(SiNode) == 32, "nanoflann node layout");

struct SiBuilder {
  const float *val;
  int dim;
  size_t leaf_max;
  std::vector<uint64_t> vind;
  // nodes of a subtree in pre-order; subtrees are built independently and spliced
  void build(size_t l, size_t r, std::vector<SiNode> &out) {
    SiNode nd;
    memset(&nd, 0, sizeof nd);
    if (r - l <= leaf_max) {
      nd.u.lr.left = l;
      nd.u.lr.right = r;
      out.push_back(nd);
      return;
    }
    // widest dimension of the points' own bounding box
    float lo[16], hi[16];
    for (int d = 0; d < dim; ++d) {
      lo[d] = 3.0e38f;
      hi[d] = -3.0e38f;
    }
    for (size_t i = l; i < r; ++i) {
      const float *v = val + vind[i];
      for (int d = 0; d < dim; ++d) {
        lo[d] = std::min(lo[d], v[d]);
        hi[d] = std::max(hi[d], v[d]);
      }
    }
    int cut = 0;
    for (int d = 1; d < dim; ++d)
      if (hi[d] - lo[d] > hi[cut] - lo[cut]) cut = d;
    const size_t mid = l + (r - l) / 2;
    std::nth_element(vind.begin() + l, vind.begin() + mid, vind.begin() + r,
                     [&](uint64_t a, uint64_t b) { return val[a + cut] < val[b + cut]; });
    float divlow = -3.0e38f;
    for (size_t i = l; i < mid; ++i) divlow = std::max(divlow, val[vind[i] + cut]);
    nd.u.sub.divfeat = cut;
    nd.u.sub.divlow = divlow;                 // max of child1
    nd.u.sub.divhigh = val[vind[mid] + cut];  // min of child2 (the nth element)
    nd.child1 = nd.child2 = 1;                // non-NULL
    const size_t at = out.size();
    out.push_back(nd);
    if (r - l >= (size_t)1 << 16) {
      // big subtrees: the two halves in parallel, spliced in pre-order
      std::vector<SiNode> left, right;
#pragma omp task shared(left) if (r - l >= (size_t)1 << 18)
      build(l, mid, left);
      build(mid, r, right);
#pragma omp taskwait
      out.insert(out.end(), left.begin(), left.end());
      out.insert(out.end(), right.begin(), right.end());
    } else {
      build(l, mid, out);
      build(mid, r, out);
    }
    (void)at;
  }
};

}  // namespace

extern "C" int smbh_si_write(const char *prefix, const float *val, size_t n_points, int dim, int max_leaf) {
  if (dim < 1 || dim > 16 || max_leaf < 1 || n_points < (size_t)dim) return SMB_ERR_ARG;
  const size_t m = n_points - (size_t)dim + 1;  // window points (sigmap_adaptor.h:89-91)
  SiBuilder b;
  b.val = val;
  b.dim = dim;
  b.leaf_max = (size_t)max_leaf;
  b.vind.resize(m);
  for (size_t i = 0; i < m; ++i) b.vind[i] = i;
  std::vector<SiNode> nodes;
#pragma omp parallel
#pragma omp single
  b.build(0, m, nodes);
  // root bounding box: exact, over all window points (computeInitialDistances uses it)
  std::vector<float> box(2 * (size_t)dim);
  for (int d = 0; d < dim; ++d) {
    float lo = 3.0e38f, hi = -3.0e38f;
    for (size_t i = 0; i < m; ++i) {
      lo = std::min(lo, val[i + d]);
      hi = std::max(hi, val[i + d]);
    }
    box[2 * d] = lo;
    box[2 * d + 1] = hi;
  }
  const std::string path = std::string(prefix) + ".si";
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return SMB_ERR_IO;
  const uint64_t m64 = m, dim64 = (uint64_t)dim, leaf64 = (uint64_t)max_leaf;
  const int32_t dim32 = dim;
  bool ok = fwrite(&m64, 8, 1, f) == 1 && fwrite(&dim32, 4, 1, f) == 1 && fwrite(&dim64, 8, 1, f) == 1 &&
            fwrite(box.data(), sizeof(float), box.size(), f) == box.size() && fwrite(&leaf64, 8, 1, f) == 1 &&
            fwrite(&m64, 8, 1, f) == 1 && fwrite(b.vind.data(), 8, m, f) == m &&
            fwrite(nodes.data(), sizeof(SiNode), nodes.size(), f) == nodes.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? SMB_OK : SMB_ERR_IO;
}
