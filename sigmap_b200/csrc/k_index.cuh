// k_index.cuh -- K4: the flat device-resident spatial index that replaces nanoflann's
// KD-tree (nanoflann.hpp:858-1004 build, :1278-1410 radius search), and its build kernels.
//
// Layout (see IndexView in sb_device.cuh): the N-5 window points value[w..w+5]
// (sigmap_adaptor.h:89-97) are sorted by a 60-bit Morton code (6 dims x 10 bits) and cut into
// 32-point leaf blocks stored SoA ([block][dim][lane]) so a warp evaluates one block with six
// coalesced 128-byte loads; axis-aligned boxes of 32 consecutive blocks / boxes form a
// pointer-free fan-out-32 hierarchy.  A query is handled by ONE WARP: it pops (level, group)
// entries from a small shared-memory stack, every lane tests one child box against the query
// ball, and the ballot decides what to push / which leaf blocks to evaluate.
//
// Exactness: the accept test is the reference's own fp32 expression
//   d2 = ((e0+e1)+e2)+e3, then +e4, +e5, e_k = (q_k - v_k)^2, accept iff d2 < radius
// (nanoflann.hpp:383-408, :249-251, :1362; the "radius" is already squared, Q4) without FMA.
// Boxes only prune, with a relative slack of 1e-4 on the squared radius, so no point the
// exact test would accept is ever lost.
#ifndef SB_K_INDEX_CUH
#define SB_K_INDEX_CUH

#include "sb_device.cuh"

namespace sb {

constexpr float kPadValue = 1.0e18f;  // padding points / boxes: never inside any ball

// ------------------------------------------------------------------ build
__device__ __forceinline__ uint64_t spread10(uint32_t v) {
  // bit i of v -> bit 6*i
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) r |= (uint64_t)((v >> i) & 1u) << (6 * i);
  return r;
}

__global__ void k_morton(const float *__restrict__ val, uint64_t n_windows, float vmin, float inv_span,
                         uint64_t *__restrict__ code, uint32_t *__restrict__ widx) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_windows) return;
  uint64_t c = 0;
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    float t = (val[w + d] - vmin) * inv_span * 1024.0f;
    int qv = (int)t;
    qv = qv < 0 ? 0 : (qv > 1023 ? 1023 : qv);
    c |= spread10((uint32_t)qv) << (kDim - 1 - d);
  }
  code[w] = c;
  widx[w] = (uint32_t)w;
}

// sorted rank i -> leaf arrays; one thread per slot of the padded leaf array
__global__ void k_build_leaves(const float *__restrict__ val, const uint64_t *__restrict__ pos,
                               const uint32_t *__restrict__ order, uint64_t n_windows,
                               uint32_t n_blocks, float *__restrict__ leaf_vals,
                               uint32_t *__restrict__ leaf_tpos, uint32_t *__restrict__ leaf_bucket,
                               uint32_t *__restrict__ leaf_widx) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n_blocks * kLeaf) return;
  const uint32_t blk = (uint32_t)(i / kLeaf), lane = (uint32_t)(i % kLeaf);
  float *v = leaf_vals + (size_t)blk * kDim * kLeaf + lane;
  if (i < n_windows) {
    const uint32_t w = order[i];
#pragma unroll
    for (int d = 0; d < kDim; ++d) v[d * kLeaf] = val[(uint64_t)w + d];
    const uint64_t P = pos[w];
    leaf_tpos[i] = (uint32_t)(P >> 1);                              // spatial_index.cc:380-381
    leaf_bucket[i] = (uint32_t)((P >> 33) << 1) | (uint32_t)(P & 1); // contig*2 + strand
    leaf_widx[i] = w;
  } else {
#pragma unroll
    for (int d = 0; d < kDim; ++d) v[d * kLeaf] = kPadValue;
    leaf_tpos[i] = 0;
    leaf_bucket[i] = 0xFFFFFFFFu;
    leaf_widx[i] = 0xFFFFFFFFu;
  }
}

// boxes of level 0: one warp per leaf block; box j of a level is stored in group j/32, lane j%32
__device__ __forceinline__ void store_box(float *level, uint32_t j, const float *lo, const float *hi) {
  float *g = level + (size_t)(j / kFan) * 12 * kFan + (j % kFan);
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    g[d * kFan] = lo[d];
    g[(kDim + d) * kFan] = hi[d];
  }
}

__global__ void k_boxes_level0(const float *__restrict__ leaf_vals, const uint32_t *__restrict__ leaf_bucket,
                               uint32_t n_blocks, uint32_t n_boxes_padded, float *__restrict__ level) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
  if (warp >= n_boxes_padded) return;
  float lo[kDim], hi[kDim];
  const bool real = warp < n_blocks && leaf_bucket[(size_t)warp * kLeaf + lane] != 0xFFFFFFFFu;
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    float v = real ? leaf_vals[((size_t)warp * kDim + d) * kLeaf + lane] : 0.0f;
    lo[d] = real ? v : kPadValue;
    hi[d] = real ? v : -kPadValue;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
  }
  if (lane == 0) store_box(level, warp, lo, hi);
}

// boxes of level l+1 from level l: one warp per parent (= one group of level l)
__global__ void k_boxes_up(const float *__restrict__ child, uint32_t n_child, uint32_t n_parent_padded,
                           float *__restrict__ parent) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
  if (warp >= n_parent_padded) return;
  float lo[kDim], hi[kDim];
  const uint32_t j = warp * kFan + lane;
  const bool real = j < n_child;
  const float *g = child + (size_t)warp * 12 * kFan + lane;
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    lo[d] = real ? g[d * kFan] : kPadValue;
    hi[d] = real ? g[(kDim + d) * kFan] : -kPadValue;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
  }
  if (lane == 0) store_box(parent, warp, lo, hi);
}

// ------------------------------------------------------------------ search
constexpr int kSearchWarps = 8;          // warps per CTA
constexpr int kStackCap = 32 * kMaxLevels;
constexpr int kStageCap = 128;           // staged hits per warp before one global reservation

struct SearchArgs {
  // queries: either pipeline mode (features of batch entries) or stage mode (explicit)
  const float *features;       // pipeline: rows of kFeatCap floats; stage: [nq][6]
  const uint32_t *feat_row;    // pipeline: feature row of each batch entry
  const uint32_t *q_off;       // pipeline: exclusive scan of queries per entry, B+1 entries
  const uint32_t *entry_slot;  // pipeline: batch entry -> slot
  const SlotState *slots;      // pipeline: num_events (query offset) per slot
  uint32_t B;
  uint32_t n_queries;          // stage mode; pipeline reads q_off[B]
  int step;
  float radius;                // squared L2 (Q4)
  KeyLayout key;
  uint64_t *out_key;
  float *out_dist;
  unsigned long long cap;      // capacity of out_key/out_dist
  Counters *ctr;
  SlotState *slots_mut;        // to flag capped queries
};

__device__ __forceinline__ float exact_d2(const float q[kDim], const float v[kDim]) {
  float e[kDim];
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    float t = __fsub_rn(q[d], v[d]);
    e[d] = __fmul_rn(t, t);
  }
  float r = __fadd_rn(__fadd_rn(__fadd_rn(e[0], e[1]), e[2]), e[3]);
  r = __fadd_rn(r, e[4]);
  r = __fadd_rn(r, e[5]);
  return r;
}

// STAGE=false: hits become sort keys (entry|bucket|target|query) + d2, capped at 5000/query.
// STAGE=true : key = query_id << 32 | window index, no cap (parity hook, compared as sets).
template <bool STAGE>
__global__ void __launch_bounds__(kSearchWarps * 32)
k_radius_search(const IndexView ix, const SearchArgs a) {
  __shared__ uint32_t s_stack[kSearchWarps][kStackCap];
  __shared__ uint64_t s_key[kSearchWarps][kStageCap];
  __shared__ float s_dist[kSearchWarps][kStageCap];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t *stack = s_stack[wid];
  uint64_t *st_key = s_key[wid];
  float *st_dist = s_dist[wid];
  const uint32_t nq = STAGE ? a.n_queries : a.q_off[a.B];
  const float r2 = a.radius;
  const float r2_prune = r2 * 1.0001f + 1e-12f;
  int staged = 0;
  unsigned long long my_hits = 0, my_capped = 0;

  auto flush = [&]() {
    if (staged == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&a.ctr->n_anchors, (unsigned long long)staged);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < staged; i += 32) {
      unsigned long long o = base + i;
      if (o < a.cap) {
        a.out_key[o] = st_key[i];
        a.out_dist[o] = st_dist[i];
      }
    }
    __syncwarp();
    staged = 0;
  };

  for (;;) {
    // dynamic work distribution: 4 queries per grab
    uint32_t q0 = 0;
    if (lane == 0) q0 = atomicAdd(&a.ctr->work, 4u);
    q0 = __shfl_sync(0xffffffffu, q0, 0);
    if (q0 >= nq) break;
    const uint32_t q1 = min(q0 + 4u, nq);
    for (uint32_t qi = q0; qi < q1; ++qi) {
      // ---- locate the query
      float q[kDim];
      uint32_t entry = 0, qpos = 0, slot = 0;
      if (STAGE) {
#pragma unroll
        for (int d = 0; d < kDim; ++d) q[d] = __ldg(a.features + (size_t)qi * kDim + d);
      } else {
        // binary search: largest entry with q_off[entry] <= qi
        uint32_t lo = 0, hi = a.B;
        while (hi - lo > 1) {
          uint32_t mid = (lo + hi) >> 1;
          if (__ldg(a.q_off + mid) <= qi) lo = mid; else hi = mid;
        }
        entry = lo;
        const uint32_t k = qi - __ldg(a.q_off + entry);
        const uint32_t p = (uint32_t)a.step * (k + 1);   // seeds at step, 2*step, ... (Q3)
        slot = __ldg(a.entry_slot + entry);
        qpos = p + a.slots[slot].num_events;              // position + query_start_offset
        const float *f = a.features + (size_t)__ldg(a.feat_row + entry) * kFeatCap + p;
#pragma unroll
        for (int d = 0; d < kDim; ++d) q[d] = __ldg(f + d);
      }
      uint32_t qhits = 0;
      bool capped = false;
      int sp = 0;
      if (lane == 0) stack[0] = ((uint32_t)(ix.n_levels - 1) << 27);  // (top level, group 0)
      sp = 1;
      __syncwarp();
      while (sp > 0) {
        const uint32_t top = stack[sp - 1];
        --sp;
        __syncwarp();
        const int level = (int)(top >> 27);
        const uint32_t group = top & 0x07FFFFFFu;
        // ---- every lane tests one child box of (level, group)
        const float *g = ix.level_box[level] + (size_t)group * 12 * kFan + lane;
        // boxes only prune (with slack), so this distance may use FMA; the accept test below
        // may not
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < kDim; ++d) {
          const float lo = __ldg(g + d * kFan), hi = __ldg(g + (kDim + d) * kFan);
          const float dd = fmaxf(fmaxf(lo - q[d], q[d] - hi), 0.0f);
          s = __fmaf_rn(dd, dd, s);
        }
        uint32_t mask = __ballot_sync(0xffffffffu, s <= r2_prune);
        if (level > 0) {
          // push child groups (level-1, group*32 + bit)
          const int n = __popc(mask);
          if (mask & (1u << lane)) {
            const int at = sp + __popc(mask & ((1u << lane) - 1u));
            stack[at] = ((uint32_t)(level - 1) << 27) | (group * kFan + lane);
          }
          sp += n;
          __syncwarp();
          continue;
        }
        // ---- level 0: surviving children are leaf blocks; evaluate them one by one
        while (mask) {
          const int bit = __ffs(mask) - 1;
          mask &= mask - 1;
          const uint32_t blk = group * kFan + bit;
          const float *lv = ix.leaf_vals + (size_t)blk * kDim * kLeaf + lane;
          float v[kDim];
#pragma unroll
          for (int d = 0; d < kDim; ++d) v[d] = __ldg(lv + d * kLeaf);
          const float d2 = exact_d2(q, v);
          uint32_t hit = __ballot_sync(0xffffffffu, d2 < r2);
          if (!hit) continue;
          if (!STAGE) {
            const uint32_t room = kMaxHits - qhits;
            if ((uint32_t)__popc(hit) > room) {
              // keep the first `room` hits in lane order (deterministic)
              uint32_t keep = 0, m = hit;
              for (uint32_t c = 0; c < room; ++c) {
                keep |= m & (0u - m);
                m &= m - 1;
              }
              hit = keep;
              capped = true;
            }
          }
          const int nh = __popc(hit);
          if (staged + nh > kStageCap) flush();
          if (hit & (1u << lane)) {
            const int at = staged + __popc(hit & ((1u << lane) - 1u));
            const size_t pi = (size_t)blk * kLeaf + lane;
            uint64_t key;
            if (STAGE) {
              key = ((uint64_t)qi << 32) | __ldg(ix.leaf_widx + pi);
            } else {
              key = a.key.pack(entry, __ldg(ix.leaf_bucket + pi), __ldg(ix.leaf_tpos + pi), qpos);
            }
            st_key[at] = key;
            st_dist[at] = d2;
          }
          __syncwarp();
          staged += nh;
          qhits += nh;
          if (!STAGE && qhits >= kMaxHits) {
            capped = true;  // conservatively: there may have been more than 5000
            sp = 0;  // the reference stops taking hits after 5000 (spatial_index.cc:371-372)
            break;
          }
        }
      }
      my_hits += qhits;
      if (capped) {
        ++my_capped;
        if (!STAGE && lane == 0) atomicOr(&a.slots_mut[slot].flags, 1u);
      }
    }
  }
  flush();
  if (lane == 0) {
    if (my_hits) atomicAdd(&a.ctr->n_hits, my_hits);
    if (my_capped) atomicAdd(&a.ctr->n_capped, my_capped);
  }
}

}  // namespace sb
#endif
