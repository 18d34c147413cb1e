// k_sort.cuh -- K5: per-entry (per read chunk) anchor sort in shared memory.
//
// The reference sorts every (strand, contig) bucket of one chunk with std::sort on
// (target, query, distance) (spatial_index.cc:411-417, key spatial_index.h:22-25).  A global
// radix sort of the whole batch moves every 12-byte anchor through HBM six times.  Here the
// search kernel already knows which chunk ("entry") a hit belongs to and records where it wrote
// each run of hits (RunRec), so the sort never has to look at the entry bits: ONE CTA per
// entry gathers the entry's runs into its 227 KB of shared memory, sorts them there and writes
// every anchor back exactly once, in (bucket, target, query) order.
//
//   pass 0   (only if the entry does not fit at once) coarse histogram of the linear coordinate
//            g = bucket_base[bucket] + target over <= 256 bins -> contiguous parts of <= kSortCap
//   per part gather (filter by g range) into shared memory + fine histogram (8192 bins)
//            -> exclusive scan -> counting-sort scatter of 16-bit slots -> each bin (1.1 anchors
//            on average) ordered by full key: insertion sort by one thread, or a warp-wide rank
//            sort for the few dense bins (the true locus, carried chains)
//            -> coalesced write-back at a position reserved with one atomic per entry.
//
// Output segments of different entries land in arrival order; inside an entry the order is the
// reference's.  Nothing downstream depends on the order of entries (k_chain_prep finds segment
// bounds from key changes).  (target, query) pairs are unique inside a bucket, so the order is
// total and the result is bit-identical to the radix-sort path.
#ifndef SB_K_SORT_CUH
#define SB_K_SORT_CUH

#include "sb_device.cuh"

namespace sb {

constexpr int kRunsCap = 512;        // runs recorded per entry (a run = one flush of <= 128 hits)
// Two shapes of the kernel (template parameters CAP = anchors of one part held in shared memory,
// THREADS, BINS = fine bins per part):
//   <10240, 1024, 8192>  one 200 KB CTA per SM: fewest passes over an entry's runs
//   < 5120,  512, 4096>  two 100 KB CTAs per SM: one CTA's barrier / DRAM waits overlap the
//                        other's work, at the price of twice the parts per entry
constexpr int kSortCapBig = 10240, kSortCapSmall = 5120;
constexpr int kCoarseBins = 256;
constexpr int kSmallBin = 24;        // bins up to this size: insertion sort by one thread
constexpr int kBigBinCap = 512;      // dense bins queued for the warp-wide rank sort

struct SegSortArgs {
  const uint64_t *key_in;
  const float *dist_in;
  uint64_t *key_out;
  float *dist_out;
  const RunRec *runs;           // [B][kRunsCap]
  const uint32_t *run_count;    // [B]
  uint32_t B;
  KeyLayout kl;
  const uint64_t *bucket_base;  // [n_buckets + 1] linear coordinate of target 0 of each bucket
  int gshift;                   // coarse bin = g >> gshift, < kCoarseBins
  uint32_t n_coarse;
  Counters *ctr;                // sort_cursor (output position), error bit 4 (dense coarse bin)
};

constexpr size_t sort_smem_bytes(int cap, int bins) {
  return (size_t)cap * (8 + 4 + 2 + 2) + (size_t)bins * 4 + (size_t)kCoarseBins * 4 +
         (size_t)(kCoarseBins + 2) * 4 + (size_t)kBigBinCap * 4 + 64 * 4;
}

template <int kSortCap, int kSortThreads, int kSortBins>
__global__ void __launch_bounds__(kSortThreads, kSortThreads == 1024 ? 1 : 2) k_seg_sort(const SegSortArgs a) {
  static_assert(kSortBins % kSortThreads == 0, "bins per thread");
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint64_t *s_key = reinterpret_cast<uint64_t *>(s_raw);
  float *s_dist = reinterpret_cast<float *>(s_key + kSortCap);
  uint32_t *s_bins = reinterpret_cast<uint32_t *>(s_dist + kSortCap);
  uint32_t *s_coarse = s_bins + kSortBins;
  uint32_t *s_part = s_coarse + kCoarseBins;      // [kCoarseBins + 2] part boundaries (coarse bins)
  uint32_t *s_big = s_part + kCoarseBins + 2;     // [kBigBinCap]
  uint32_t *s_misc = s_big + kBigBinCap;          // [64]: 0..31 warp sums, 32 n, 33 nbig, 34 np, 35 total, 36/37 base
  uint16_t *s_order = reinterpret_cast<uint16_t *>(s_misc + 64);
  uint16_t *s_order2 = s_order + kSortCap;

  const uint32_t entry = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned full = 0xffffffffu;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t nr = min(a.run_count[entry], (uint32_t)kRunsCap);
  if (nr == 0) return;
  const RunRec *runs = a.runs + (size_t)entry * kRunsCap;
  const KeyLayout kl = a.kl;

  // ---- total anchors of the entry
  uint32_t mine = 0;
  for (uint32_t r = tid; r < nr; r += kSortThreads) mine += runs[r].count;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(full, mine, d);
  if (tid < kCoarseBins) s_coarse[tid] = 0;
  if (tid < 32) s_misc[tid] = 0;
  __syncthreads();
  if (lane == 0) s_misc[wid] = mine;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = s_misc[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(full, v, d);
    if (lane == 0) s_misc[35] = v;
  }
  __syncthreads();
  const uint32_t total = s_misc[35];
  if (total == 0) return;

  auto coord = [&](uint64_t k) -> uint64_t { return __ldg(a.bucket_base + kl.bucket(k)) + kl.target(k); };

  // ---- pass 0: coarse histogram and the split into parts that fit shared memory
  if (total > (uint32_t)kSortCap) {
    for (uint32_t r = wid; r < nr; r += kSortThreads / 32) {
      const RunRec run = runs[r];
      for (uint32_t i0 = 0; i0 < run.count; i0 += 4 * 32) {  // a run is <= 128 hits: one trip
        uint64_t kq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = i0 + u * 32 + lane;
          kq[u] = i < run.count ? a.key_in[run.start + i] : ~0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (kq[u] != ~0ull) atomicAdd(&s_coarse[(uint32_t)(coord(kq[u]) >> a.gshift)], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t np = 0, acc = 0;
      bool dense = false;
      s_part[0] = 0;
      for (uint32_t b = 0; b < a.n_coarse; ++b) {
        const uint32_t cb = s_coarse[b];
        if (cb > (uint32_t)kSortCap) dense = true;
        if (acc + cb > (uint32_t)kSortCap) {
          s_part[++np] = b;
          acc = 0;
        }
        acc += cb;
      }
      s_part[++np] = a.n_coarse;
      s_misc[34] = dense ? 0u : np;
      if (dense) atomicOr(&a.ctr->error, 16u);  // the caller falls back to the global sort
    }
  } else if (tid == 0) {
    s_part[0] = 0;
    s_part[1] = a.n_coarse;
    s_misc[34] = 1;
  }
  if (tid == 0) {
    const unsigned long long base = atomicAdd(&a.ctr->sort_cursor, (unsigned long long)total);
    s_misc[36] = (uint32_t)base;
    s_misc[37] = (uint32_t)(base >> 32);
  }
  __syncthreads();
  const uint32_t np = s_misc[34];
  unsigned long long out = ((unsigned long long)s_misc[37] << 32) | s_misc[36];

  for (uint32_t p = 0; p < np; ++p) {
    const uint32_t c_lo = s_part[p], c_hi = s_part[p + 1];
    const uint64_t g_lo = (uint64_t)c_lo << a.gshift;
    const uint64_t span = (uint64_t)(c_hi - c_lo) << a.gshift;
    int fshift = 0;
    while ((span >> fshift) >= (uint64_t)kSortBins) ++fshift;
    for (int b = tid; b < kSortBins; b += kSortThreads) s_bins[b] = 0;
    if (tid == 0) {
      s_misc[32] = 0;
      s_misc[33] = 0;
    }
    __syncthreads();

    // ---- gather the part's anchors + fine histogram
    for (uint32_t r = wid; r < nr; r += kSortThreads / 32) {
      const RunRec run = runs[r];
      for (uint32_t i0 = 0; i0 < run.count; i0 += 4 * 32) {  // a run is <= 128 hits: one trip,
        uint64_t kq[4];                                       // its eight loads in flight together
        float dq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = i0 + u * 32 + lane;
          const bool has = i < run.count;
          kq[u] = has ? a.key_in[run.start + i] : ~0ull;
          dq[u] = has ? a.dist_in[run.start + i] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + u * 32 >= run.count) break;
          const uint64_t k = kq[u];
          uint64_t g = 0;
          bool in = false;
          if (k != ~0ull) {
            g = coord(k);
            const uint32_t cb = (uint32_t)(g >> a.gshift);
            in = cb >= c_lo && cb < c_hi;
          }
          const unsigned m = __ballot_sync(full, in);
          if (!m) continue;
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&s_misc[32], (uint32_t)__popc(m));
          base = __shfl_sync(full, base, 0);
          if (in) {
            const uint32_t slot = base + __popc(m & lt);
            if (slot < (uint32_t)kSortCap) {  // always true: part sizes are exact
              s_key[slot] = k;
              s_dist[slot] = dq[u];
              atomicAdd(&s_bins[(uint32_t)((g - g_lo) >> fshift)], 1u);
            }
          }
        }
      }
    }
    __syncthreads();
    const uint32_t n = min(s_misc[32], (uint32_t)kSortCap);

    // ---- exclusive scan of the fine histogram (8 consecutive bins per thread)
    {
      uint32_t v[kSortBins / kSortThreads], sum = 0;
#pragma unroll
      for (int j = 0; j < kSortBins / kSortThreads; ++j) {
        v[j] = s_bins[tid * (kSortBins / kSortThreads) + j];
        sum += v[j];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(full, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_misc[wid] = incl;
      __syncthreads();
      if (wid == 0) {
        const uint32_t w = lane < kSortThreads / 32 ? s_misc[lane] : 0u;
        uint32_t in2 = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(full, in2, d);
          if (lane >= d) in2 += t;
        }
        s_misc[lane] = in2 - w;
      }
      __syncthreads();
      uint32_t run_sum = s_misc[wid] + incl - sum;
#pragma unroll
      for (int j = 0; j < kSortBins / kSortThreads; ++j) {
        s_bins[tid * (kSortBins / kSortThreads) + j] = run_sum;
        run_sum += v[j];
      }
    }
    __syncthreads();

    // ---- counting-sort scatter of the slots; afterwards s_bins[b] = end of bin b
    for (uint32_t slot = tid; slot < n; slot += kSortThreads) {
      const uint32_t bin = (uint32_t)((coord(s_key[slot]) - g_lo) >> fshift);
      s_order[atomicAdd(&s_bins[bin], 1u)] = (uint16_t)slot;
    }
    __syncthreads();

    // ---- order every bin by the full key (bucket | target | query)
    for (int b = tid; b < kSortBins; b += kSortThreads) {
      const uint32_t lo = b ? s_bins[b - 1] : 0u, hi = s_bins[b];
      const uint32_t m = hi - lo;
      if (m < 2) continue;
      bool by_thread = m <= (uint32_t)kSmallBin;
      if (!by_thread) {
        const uint32_t at = atomicAdd(&s_misc[33], 1u);
        if (at < (uint32_t)kBigBinCap) s_big[at] = (uint32_t)b;
        else by_thread = true;  // queue full: slow but correct
      }
      if (by_thread) {
        for (uint32_t x = lo + 1; x < hi; ++x) {
          const uint16_t ox = s_order[x];
          const uint64_t kx = s_key[ox];
          uint32_t y = x;
          while (y > lo && s_key[s_order[y - 1]] > kx) {
            s_order[y] = s_order[y - 1];
            --y;
          }
          s_order[y] = ox;
        }
      }
    }
    __syncthreads();
    const uint32_t nbig = min(s_misc[33], (uint32_t)kBigBinCap);
    for (uint32_t bi = wid; bi < nbig; bi += kSortThreads / 32) {
      const uint32_t b = s_big[bi];
      const uint32_t lo = b ? s_bins[b - 1] : 0u, hi = s_bins[b];
      const uint32_t m = hi - lo;
      for (uint32_t x = lane; x < m; x += 32) {
        const uint16_t ox = s_order[lo + x];
        const uint64_t kx = s_key[ox];
        uint32_t rank = 0;
        for (uint32_t y = 0; y < m; ++y) {
          const uint64_t ky = s_key[s_order[lo + y]];
          rank += (ky < kx || (ky == kx && y < x)) ? 1u : 0u;
        }
        s_order2[lo + rank] = ox;
      }
      __syncwarp();
      for (uint32_t x = lane; x < m; x += 32) s_order[lo + x] = s_order2[lo + x];
    }
    __syncthreads();

    // ---- write the part back, sorted
    for (uint32_t pos = tid; pos < n; pos += kSortThreads) {
      const uint32_t slot = s_order[pos];
      a.key_out[out + pos] = s_key[slot];
      a.dist_out[out + pos] = s_dist[slot];
    }
    out += n;
    __syncthreads();
  }
}

}  // namespace sb
#endif
