// host_sim.cc -- synthetic inputs for tests and benchmarks: an i.i.d.
// uniform ACGT reference and R9.4-like raw reads simulated from the 6-mer pore model.
// Every read is generated from its own RNG streams keyed by (seed, read index), so any
// rank can generate any contiguous range of reads and get identical bytes.
//
// Read model: uniform contig (by length), uniform start, uniform strand, length
// U[min_bases, max_bases) bases; per 6-mer (true lookup, no reference quirk) a dwell of
// max(1, round(Gamma(k=2, theta=4000/450/2))) samples; sample pA = level_mean +
// N(0,1) * level_stdv * noise; raw = round(pA * 8192 / 1437.976685 - 10) as int16
// (digitisation 8192, range 1437.976685, offset 10).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/sigmap_b200.h"
#include "sb_host.h"

namespace {
constexpr double kDigitisation = 8192.0, kRange = 1437.976685, kOffset = 10.0;
constexpr double kTheta = 4000.0 / 450.0 / 2.0;

struct ReadPlan {
  uint32_t contig, start, len, strand_plus;
};

ReadPlan plan_read(sb::Rng &r, const uint32_t *lengths, uint32_t n_contigs, uint64_t total,
                   uint32_t min_bases, uint32_t max_bases) {
  ReadPlan p;
  p.len = min_bases + (uint32_t)r.below(max_bases - min_bases);
  // contig proportional to length; retry until the read fits
  for (;;) {
    uint64_t x = r.below(total);
    uint32_t c = 0;
    while (c + 1 < n_contigs && x >= lengths[c]) x -= lengths[c++];
    if (lengths[c] > p.len) {
      p.contig = c;
      p.start = (uint32_t)r.below(lengths[c] - p.len);
      break;
    }
    if (p.len > min_bases) p.len = min_bases + (p.len - min_bases) / 2;
  }
  p.strand_plus = (uint32_t)(r.next() >> 63);
  return p;
}

inline uint32_t dwell(sb::Rng &r) {
  double g = -kTheta * (std::log(r.uniform_pos()) + std::log(r.uniform_pos()));  // Gamma(2, theta)
  long d = std::lround(g);
  return d < 1 ? 1u : (uint32_t)d;
}
}  // namespace

extern "C" {

int smbh_sim_reference(uint64_t seed, const uint32_t *lengths, uint32_t n_contigs, char **seqs) {
  for (uint32_t c = 0; c < n_contigs; ++c) {
    sb::Rng r(seed, 0x5EF00000ull + c);
    char *s = seqs[c];
    uint32_t i = 0;
    while (i < lengths[c]) {
      uint64_t w = r.next();
      for (int k = 0; k < 32 && i < lengths[c]; ++k, w >>= 2) s[i++] = "ACGT"[w & 3];
    }
    s[lengths[c]] = 0;
  }
  return SMB_OK;
}

int smbh_sim_reads(uint64_t seed, const char *const *seqs, const uint32_t *lengths,
                   uint32_t n_contigs, const float *level_mean, const float *level_stdv,
                   uint64_t first_read, uint64_t n_reads, uint32_t min_bases,
                   uint32_t max_bases, float noise, uint64_t *read_off, int16_t *raw,
                   uint32_t *truth) {
  if (max_bases <= min_bases || min_bases < 6) return SMB_ERR_ARG;
  uint64_t total = 0;
  for (uint32_t c = 0; c < n_contigs; ++c) total += lengths[c];
  const bool fill = raw != nullptr;
  if (!fill) read_off[0] = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (uint64_t i = 0; i < n_reads; ++i) {
    const uint64_t id = first_read + i;
    sb::Rng rs(seed, 2 * id + 1);  // structure stream: placement + dwells
    ReadPlan p = plan_read(rs, lengths, n_contigs, total, min_bases, max_bases);
    const uint32_t n_kmers = p.len - 5;
    if (!fill) {
      uint64_t ns = 0;
      for (uint32_t k = 0; k < n_kmers; ++k) ns += dwell(rs);
      read_off[i + 1] = ns;  // per-read count for now; prefix-summed below
      if (truth) {
        truth[4 * i + 0] = p.contig;
        truth[4 * i + 1] = p.start;
        truth[4 * i + 2] = p.start + p.len;
        truth[4 * i + 3] = p.strand_plus;
      }
      continue;
    }
    sb::Rng rn(seed, 2 * id + 2);  // noise stream
    const char *ref = seqs[p.contig];
    int16_t *out = raw + read_off[i];
    uint32_t h = 0;
    auto base_at = [&](uint32_t k) -> int {  // k-th base of the read in read orientation
      int c = sb::base_code(p.strand_plus ? ref[p.start + k] : ref[p.start + p.len - 1 - k]);
      if (c < 0) c = 0;  // ambiguous reference base: simulate as 'A'
      return p.strand_plus ? c : 3 ^ c;
    };
    for (uint32_t k = 0; k < 5; ++k) h = (h << 2) | (uint32_t)base_at(k);
    uint64_t w = 0;
    bool have_spare = false;
    double spare = 0;
    for (uint32_t k = 0; k < n_kmers; ++k) {
      h = ((h << 2) | (uint32_t)base_at(k + 5)) & 4095u;
      const double mu = level_mean[h], sd = (double)level_stdv[h] * noise;
      const uint32_t d = dwell(rs);
      for (uint32_t j = 0; j < d; ++j) {
        double z;  // Box-Muller, both outputs used
        if (have_spare) {
          z = spare;
          have_spare = false;
        } else {
          double u1 = rn.uniform_pos(), u2 = rn.uniform();
          double m = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586 * u2;
          z = m * std::cos(a);
          spare = m * std::sin(a);
          have_spare = true;
        }
        double pa = mu + z * sd;
        long v = std::lround(pa * kDigitisation / kRange - kOffset);
        if (v < -32768) v = -32768;
        if (v > 32767) v = 32767;
        out[w++] = (int16_t)v;
      }
    }
  }
  if (!fill)
    for (uint64_t i = 0; i < n_reads; ++i) read_off[i + 1] += read_off[i];
  return SMB_OK;
}

}  // extern "C"
