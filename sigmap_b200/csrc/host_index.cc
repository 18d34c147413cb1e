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
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sigmap_b200.h"
#include "sb_host.h"

namespace {

constexpr int kK = 6;                    // pore-model k
constexpr int kMaskK = SMB_DIM + kK - 1; // 11-mers are masked (sigmap.cc:1014)
constexpr float kMaskFreq = 0.0002f;
constexpr double kMinDelta = 0.01;       // spatial_index.cc:46

struct StrandView {
  std::string owned;  // reverse complement when needed
  const char *s;
  uint32_t len;
};

StrandView make_strand(const char *seq, uint32_t len, int strand) {
  StrandView v;
  if (strand == 0) {
    v.s = seq;
    v.len = len;
    return v;
  }
  v.owned.resize(len);
  for (uint32_t i = 0; i < len; ++i) {  // sequence_batch.h:66-77
    int c = sb::base_code(seq[len - 1 - i]);
    v.owned[i] = c < 0 ? 'N' : "ACGT"[3 ^ c];
  }
  v.s = v.owned.c_str();  // NUL-terminated like std::string::data() in the reference
  v.len = len;
  return v;
}

// Expected current per position.  Position 0 hashes bases 0..5; position p >= 1 shifts in
// base p+6 (not p+5), so the last position shifts in the terminating NUL, read as 'A'.
void expected_levels(const StrandView &v, const float *level_mean, std::vector<float> &out) {
  const uint32_t L = v.len - kK + 1, mask = (1u << (2 * kK)) - 1;
  out.resize(L);
  uint32_t h = 0;
  for (uint32_t i = 0; i < (uint32_t)kK; ++i) {
    int c = i < L ? sb::base_code(v.s[i]) : -1;
    h = ((h << 2) | (uint32_t)(c < 0 ? 0 : c)) & mask;
  }
  out[0] = level_mean[h];
  for (uint32_t p = 1; p < L; ++p) {
    int c = sb::base_code(v.s[p + kK]);
    h = ((h << 2) | (uint32_t)(c < 0 ? 0 : c)) & mask;
    out[p] = level_mean[h];
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
  for (size_t i = 0; i < n; ++i) x[i] = (float)((x[i] - mean) / sd);
}

// Walks a strand's k-mers; fn(pos_of_kmer_start, canonical_kmer) for complete k-mers and
// fn_amb(pos) where the window *ending* at an ambiguous base starts.
template <class F, class G>
void walk_kmers(const StrandView &v, F fn, G fn_amb) {
  const uint64_t m = (1ull << (2 * kMaskK)) - 1, shift = 2ull * (kMaskK - 1);
  uint64_t fw = 0, rv = 0;
  int run = 0;
  for (uint32_t p = 0; p < v.len; ++p) {
    int c = sb::base_code(v.s[p]);
    if (c >= 0) {
      fw = ((fw << 2) | (uint64_t)c) & m;
      rv = (rv >> 2) | ((uint64_t)(3 ^ c) << shift);
      if (++run >= kMaskK) fn(p + 1 - kMaskK, fw < rv ? fw : rv);
    } else {
      run = 0;
      fw = rv = 0;
      if (p >= (uint32_t)kMaskK - 1) fn_amb(p + 1 - kMaskK);
    }
  }
}

}  // namespace

extern "C" size_t smbh_build_point_cloud(const char *const *seqs, const uint32_t *lengths,
                                         uint32_t n, const float *level_mean, uint64_t *pos,
                                         float *val) {
  // k-mer census over the + strands (canonical, so the - strand adds nothing new)
  std::vector<uint32_t> hist((size_t)1 << (2 * kMaskK), 0);
  uint64_t total = 0;
  for (uint32_t s = 0; s < n; ++s) {
    StrandView v = make_strand(seqs[s], lengths[s], 0);
    walk_kmers(v, [&](uint32_t, uint64_t k) { ++hist[k]; ++total; }, [](uint32_t) {});
  }
  size_t count = 0;
  bool any = false;
  float last = 0;
  std::vector<float> z;
  std::vector<unsigned char> masked;
  for (int strand = 0; strand < 2; ++strand) {
    for (uint32_t s = 0; s < n; ++s) {
      if (lengths[s] < (uint32_t)kMaskK) continue;
      StrandView v = make_strand(seqs[s], lengths[s], strand);
      expected_levels(v, level_mean, z);
      znormalise(z);
      masked.assign(v.len - kMaskK + 1, 0);
      walk_kmers(
          v,
          [&](uint32_t p, uint64_t k) { masked[p] = ((float)hist[k] / (float)total) > kMaskFreq; },
          [&](uint32_t p) { masked[p] = 1; });
      const uint32_t n_windows = (uint32_t)z.size() - SMB_DIM + 1;
      for (uint32_t p = 0; p < n_windows; ++p) {
        if (masked[p]) continue;
        if (p == 0 || !any || std::fabs((double)(z[p] - last)) > kMinDelta) {
          if (pos) {
            pos[count] = ((((uint64_t)s << 32) | p) << 1) | (uint64_t)strand;
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
