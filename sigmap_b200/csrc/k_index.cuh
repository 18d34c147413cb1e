// k_index.cuh -- K4: the flat device-resident spatial index that replaces nanoflann's
// KD-tree (nanoflann.hpp:858-1004 build, :1278-1410 radius search), and its build kernels.
//
// Layout (see IndexView in sb_device.cuh): the N-5 window points value[w..w+5]
// (sigmap_adaptor.h:89-97) are sorted by a 60-bit Morton code (6 dims x 10 bits) and cut into
// 8-point leaves; 8 consecutive leaves / nodes form the next level's node, pointer-free.  A
// query is handled by ONE WARP that advances FOUR tree nodes (or four leaves) per step, one per
// 8-lane group: every lane tests one child box (three 16-byte loads) or evaluates one point
// (three 8-byte loads), so each step keeps four independent 128/64-byte-coalesced request
// groups in flight and the whole warp stays busy at fan-out 8 -- where a 6-D hierarchy prunes
// far better than at fan-out 32 (the ball of radius 0.28 meets ~30 8-point leaves but ~28
// 32-point blocks: 4x fewer points to evaluate, 2.5x fewer boxes to test).
//
// Exactness: the accept test is the reference's own fp32 expression
//   d2 = ((e0+e1)+e2)+e3, then +e4, +e5, e_k = (q_k - v_k)^2, accept iff d2 < radius
// (nanoflann.hpp:383-408, :249-251, :1362; the "radius" is already squared, Q4) without FMA.
// Boxes only prune: half extents are rounded outwards when built and the test keeps a relative
// slack of 1e-4 on the squared radius, so no point the exact test would accept is ever lost.
#ifndef SB_K_INDEX_CUH
#define SB_K_INDEX_CUH

#include <cuda_fp16.h>

#include "sb_device.cuh"

namespace sb {

constexpr float kPadValue = 1.0e18f;  // padding points: never inside any ball
constexpr float kPadExtent = -1.0e18f;  // padding boxes: negative half extent, never met

// ------------------------------------------------------------------ build
__device__ __forceinline__ uint64_t spread10(uint32_t v) {
  // bit i of v -> bit 6*i
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) r |= (uint64_t)((v >> i) & 1u) << (6 * i);
  return r;
}

// wsrc (contig-sharded index): offset of local window w's first value in val[]; nullptr = w
__global__ void k_morton(const float *__restrict__ val, uint64_t n_windows, float vmin, float inv_span,
                         uint64_t *__restrict__ code, uint32_t *__restrict__ widx,
                         const uint32_t *__restrict__ wsrc) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_windows) return;
  const uint64_t v0 = wsrc ? (uint64_t)wsrc[w] : w;
  uint64_t c = 0;
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    float t = (val[v0 + d] - vmin) * inv_span * 1024.0f;
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
                               uint32_t n_leaves, float2 *__restrict__ leaf_vals,
                               uint2 *__restrict__ leaf_tb, uint32_t *__restrict__ leaf_widx,
                               const uint32_t *__restrict__ wsrc, const uint32_t *__restrict__ worig) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n_leaves * kLeaf) return;
  const uint32_t leaf = (uint32_t)(i / kLeaf), sub = (uint32_t)(i % kLeaf);
  float2 *v = leaf_vals + (size_t)leaf * 3 * kLeaf + sub;
  if (i < n_windows) {
    const uint32_t w = order[i];
    const uint64_t v0 = wsrc ? (uint64_t)wsrc[w] : (uint64_t)w;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k * kLeaf] = make_float2(val[v0 + 2 * k], val[v0 + 2 * k + 1]);
    const uint64_t P = pos[w];
    // target = pos >> 1 (spatial_index.cc:380-381); bucket = contig*2 + strand
    leaf_tb[i] = make_uint2((uint32_t)(P >> 1), (uint32_t)((P >> 33) << 1) | (uint32_t)(P & 1));
    leaf_widx[i] = worig ? worig[w] : w;
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k * kLeaf] = make_float2(kPadValue, kPadValue);
    leaf_tb[i] = make_uint2(0u, 0xFFFFFFFFu);
    leaf_widx[i] = 0xFFFFFFFFu;
  }
}

// box [lo, hi] -> twelve binary16 numbers, lo rounded down and hi rounded up, so the stored box
// always contains the fp32 one (values beyond the binary16 range become +-inf: never pruned);
// stored as child `j` of node record `rec`.  A padding child is the empty box (+inf, -inf).
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return make_float2(__half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))),
                     __half2float(__ushort_as_half((unsigned short)(v >> 16))));
}
__device__ __forceinline__ void store_child_box(uint2 *rec, int j, const float *lo, const float *hi, bool real) {
  __half l[kDim], h[kDim];
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    if (real) {
      l[d] = __float2half_rd(lo[d]);
      h[d] = __float2half_ru(hi[d]);
    } else {
      l[d] = __ushort_as_half((unsigned short)0x7C00);  // +inf
      h[d] = __ushort_as_half((unsigned short)0xFC00);  // -inf
    }
  }
  rec[j] = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
  rec[kFan + j] = make_uint2(pack_h2(l[4], l[5]), pack_h2(h[0], h[1]));
  rec[2 * kFan + j] = make_uint2(pack_h2(h[2], h[3]), pack_h2(h[4], h[5]));
}
// child `j` of a node record -> lo[6], hi[6]
__device__ __forceinline__ void load_child_box(const uint2 *rec, int j, float *lo, float *hi) {
  const uint2 a = rec[j], b = rec[kFan + j], c = rec[2 * kFan + j];
  const float2 l01 = unpack_h2(a.x), l23 = unpack_h2(a.y), l45 = unpack_h2(b.x);
  const float2 h01 = unpack_h2(b.y), h23 = unpack_h2(c.x), h45 = unpack_h2(c.y);
  lo[0] = l01.x; lo[1] = l01.y; lo[2] = l23.x; lo[3] = l23.y; lo[4] = l45.x; lo[5] = l45.y;
  hi[0] = h01.x; hi[1] = h01.y; hi[2] = h23.x; hi[3] = h23.y; hi[4] = h45.x; hi[5] = h45.y;
}

// level-0 nodes: thread (n, j) boxes leaf 8n + j
__global__ void k_nodes_level0(const float2 *__restrict__ leaf_vals, const uint2 *__restrict__ leaf_tb,
                               uint32_t n_leaves, uint32_t n_nodes, uint2 *__restrict__ nodes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * kFan) return;
  const uint32_t n = t / kFan, j = t % kFan, leaf = t;
  float lo[kDim], hi[kDim];
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    lo[d] = 3.0e38f;
    hi[d] = -3.0e38f;
  }
  bool real = false;
  if (leaf < n_leaves) {
    for (int p = 0; p < kLeaf; ++p) {
      if (leaf_tb[(size_t)leaf * kLeaf + p].y == 0xFFFFFFFFu) continue;
      real = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float2 v = leaf_vals[((size_t)leaf * 3 + k) * kLeaf + p];
        lo[2 * k] = fminf(lo[2 * k], v.x);
        hi[2 * k] = fmaxf(hi[2 * k], v.x);
        lo[2 * k + 1] = fminf(lo[2 * k + 1], v.y);
        hi[2 * k + 1] = fmaxf(hi[2 * k + 1], v.y);
      }
    }
  }
  store_child_box(nodes + (size_t)n * 3 * kFan, (int)j, lo, hi, real);
}

// level l+1 from level l: thread (n, j) boxes child node 8n + j of the level below
__global__ void k_nodes_up(const uint2 *__restrict__ child, uint32_t n_child, uint32_t n_nodes,
                           uint2 *__restrict__ nodes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * kFan) return;
  const uint32_t n = t / kFan, j = t % kFan, m = t;
  float lo[kDim], hi[kDim];
#pragma unroll
  for (int d = 0; d < kDim; ++d) {
    lo[d] = 3.0e38f;
    hi[d] = -3.0e38f;
  }
  bool real = false;
  if (m < n_child) {
    const uint2 *rec = child + (size_t)m * 3 * kFan;
    for (int p = 0; p < kFan; ++p) {
      float cl[kDim], ch[kDim];
      load_child_box(rec, p, cl, ch);
      if (cl[0] > ch[0]) continue;  // padding child (empty box)
      real = true;
#pragma unroll
      for (int d = 0; d < kDim; ++d) {
        lo[d] = fminf(lo[d], cl[d]);
        hi[d] = fmaxf(hi[d], ch[d]);
      }
    }
  }
  store_child_box(nodes + (size_t)n * 3 * kFan, (int)j, lo, hi, real);
}

// ------------------------------------------------------------------ search
constexpr int kSearchWarps = 8;          // warps per CTA
constexpr int kLevelCap = 128;           // entries per level stack: 64 waiting + the children of one 8-node step
constexpr int kLeafQueueCap = 128;       // < 8 left over + 64 pushed per step
constexpr int kStageCap = 128;           // staged hits per warp before one global reservation
constexpr int kSearchGrab = 8;           // queries per grab of the dynamic work counter
constexpr int kMaxParts = 32;            // parts per entry the flush can route to (k_part_sort)

// dynamic shared memory per warp: staging (key part, point id, d2), leaf queue, level counts,
// and one 64-entry stack per node level
__host__ __device__ inline size_t search_smem_per_warp(int n_levels) {
  return (size_t)kStageCap * 16 + (size_t)kLeafQueueCap * 4 + 16 * 4 + kMaxParts * 4 +
         (size_t)n_levels * kLevelCap * 4;
}

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
  // pipeline mode: where each entry's hits went (k_sort.cuh); runs == nullptr disables it.
  // Every flush routes its hits to the n_parts coordinate ranges ("parts": part_of(g) with
  // g = bucket_base[bucket] + target) of the entry and records one run per part it touched.
  RunRec *runs;                // [B * n_parts][runs_cap]
  uint32_t *run_count;         // [B * n_parts]
  uint32_t *entry_total;       // [B * n_parts] anchors per (entry, part) so far
  uint32_t runs_cap;
  uint32_t n_parts;            // 1..kMaxParts
  float inv_span;              // 1 / coordinates per part
  const uint64_t *bucket_base;
  uint32_t grab;               // queries per grab of the work counter (0 = kSearchGrab)
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

// largest entry e in [0, B) with q_off[e] <= qi (q_off is non-decreasing, q_off[0] = 0 <= qi
// < q_off[B]); 32-ary search by the whole warp: three rounds for 32 K entries
__device__ __forceinline__ uint32_t find_entry(const uint32_t *__restrict__ q_off, uint32_t B, uint32_t qi,
                                               int lane) {
  uint32_t lo = 0, hi = B;
  while (hi - lo > 1) {
    const uint32_t stride = (hi - lo + 31u) / 32u;
    const uint32_t idx = lo + stride * (uint32_t)lane;
    const bool le = idx < hi && __ldg(q_off + idx) <= qi;
    const int k = __popc(__ballot_sync(0xffffffffu, le)) - 1;  // lane 0 (idx = lo) is always true
    lo += stride * (uint32_t)k;
    hi = min(lo + stride, hi);
  }
  return lo;
}

// STAGE=false: hits become sort keys (entry|bucket|target|query) + d2, capped at 5000/query.
// STAGE=true : key = query_id << 32 | window index, no cap (parity hook, compared as sets).
//
// Traversal: depth-first over LEVELS, eight nodes wide.  Every level has its own small stack;
// the warp always works on the lowest non-empty level, pops up to eight of its nodes (two per
// 8-lane group), tests their 64 child boxes with six independent 16-byte loads per lane in
// flight, and pushes the survivors onto the (empty) stack of the level below -- so a level
// never holds more than the 64 children of one step.  Surviving leaves queue up and are
// evaluated eight at a time the same way.
//
// BFS = true changes the order only: the warp works on the HIGHEST level with waiting nodes as long
// as the stack below has room for the 64 children of a step (else on the lowest one, whose child
// stack is empty), so every level is consumed in full groups of eight nodes with at most one
// partial group per level -- measured on the 4.6 Mbp index: 17.9 steps per query instead of 21.1.
//
// BOX16 = true (experimental, SMB_BOX=half): the box test itself runs in packed binary16 -- two
// dimensions per instruction, no unpacking of the stored corners.  It stays conservative through a
// per-query threshold: with qh = round16(q) (|qh - q| <= 2^-11 |q|) every per-dimension distance
// computed in binary16 is at most (T_d + 2^-11 |q_d|)(1 + 2^-11) where T_d is the exact distance
// to the stored box, so a box within r of q gives a sum of squares of at most
// (r + 2^-11 |q|_2)^2 (1 + 2^-11)^8; anything above that is pruned.  Overflow gives +inf (pruned,
// correctly: such a box is farther than 65504 - |q|), inf - inf gives NaN, which max() drops.
template <bool STAGE, bool BFS = false, bool BOX16 = false>
__global__ void __launch_bounds__(kSearchWarps * 32, 4)
k_radius_search(const IndexView ix, const SearchArgs a) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int grp = lane >> 3, sub = lane & 7;
  const unsigned full = 0xffffffffu;
  const unsigned lt = (1u << lane) - 1u;
  const int n_levels = ix.n_levels;
  unsigned char *mine = s_dyn + (size_t)wid * search_smem_per_warp(n_levels);
  uint64_t *st_qk = reinterpret_cast<uint64_t *>(mine);
  uint32_t *st_pid = reinterpret_cast<uint32_t *>(mine + kStageCap * 8);
  float *st_dist = reinterpret_cast<float *>(mine + kStageCap * 12);
  uint32_t *leafq = reinterpret_cast<uint32_t *>(mine + kStageCap * 16);
  uint32_t *lcnt = leafq + kLeafQueueCap;          // [16] nodes waiting per level
  uint32_t *pcnt = lcnt + 16;                      // [kMaxParts] hits per part of the flush in progress
  uint32_t *lstk = pcnt + kMaxParts;               // [n_levels][kLevelCap]
  const uint32_t nq = STAGE ? a.n_queries : a.q_off[a.B];
  const float r2 = a.radius;
  const float r2_prune = r2 * 1.0001f + 1e-12f;
  const uint32_t grab = a.grab ? a.grab : (uint32_t)kSearchGrab;
  const int top_level = n_levels - 1;
  const uint32_t n_top = ix.level_count[top_level];  // <= 8
  int staged = 0;
  unsigned long long my_hits = 0, my_capped = 0;
  // entry of the previous query and its query range (consecutive queries mostly share it)
  uint32_t entry = 0, e_q0 = 1, e_q1 = 0, slot = 0, ev_off = 0, frow = 0;
  uint32_t staged_entry = 0;  // the entry every staged hit belongs to (a run never mixes entries)

  // staged hits -> global: one reservation, then target/bucket of every hit fetched with all
  // loads of a pass in flight (they are off the traversal's critical path here)
  auto flush = [&]() {
    if (staged == 0) return;
    if (STAGE || !a.runs) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&a.ctr->n_anchors, (unsigned long long)staged);
      base = __shfl_sync(full, base, 0);
#pragma unroll
      for (int i0 = 0; i0 < kStageCap; i0 += 32) {
        const int i = i0 + lane;
        if (i < staged) {
          const unsigned long long o = base + i;
          const uint32_t pid = st_pid[i];
          uint64_t key;
          if (STAGE) {
            key = st_qk[i] | __ldg(ix.leaf_widx + pid);
          } else {
            const uint2 tb = __ldg(ix.leaf_tb + pid);
            key = st_qk[i] | ((uint64_t)tb.y << a.key.sh_b()) | ((uint64_t)tb.x << a.key.sh_t());
          }
          if (o < a.cap) {
            a.out_key[o] = key;
            a.out_dist[o] = st_dist[i];
          }
        }
      }
      __syncwarp();
      staged = 0;
      return;
    }
    // all staged hits belong to staged_entry; route them to its parts
    pcnt[lane] = 0;
    __syncwarp();
    uint64_t key[kStageCap / 32];
    uint32_t where[kStageCap / 32];  // part << 8 | rank inside the part (this flush)
#pragma unroll
    for (int u = 0; u < kStageCap / 32; ++u) {
      const int i = u * 32 + lane;
      key[u] = 0;
      where[u] = 0;
      if (i < staged) {
        const uint2 tb = __ldg(ix.leaf_tb + st_pid[i]);
        key[u] = st_qk[i] | ((uint64_t)tb.y << a.key.sh_b()) | ((uint64_t)tb.x << a.key.sh_t());
        const uint32_t part = part_of(__ldg(a.bucket_base + tb.y) + tb.x, a.inv_span, a.n_parts);
        where[u] = (part << 8) | atomicAdd(&pcnt[part], 1u);
      }
    }
    __syncwarp();
    const uint32_t mine = pcnt[lane];  // lane p: hits of part p
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(full, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - mine;
    // both reservations are issued before either result is used: one round trip, not two
    const size_t list = (size_t)staged_entry * a.n_parts + lane;
    uint32_t r = 0;
    if (mine) r = atomicAdd(&a.run_count[list], 1u);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&a.ctr->n_anchors, (unsigned long long)staged);
    base = __shfl_sync(full, base, 0);
    if (mine) {
      if (r < a.runs_cap) a.runs[list * a.runs_cap + r] = RunRec{(uint32_t)(base + excl), mine};
      else atomicOr(&a.ctr->error, 8u);
      atomicAdd(&a.entry_total[list], mine);
    }
#pragma unroll
    for (int u = 0; u < kStageCap / 32; ++u) {
      const int i = u * 32 + lane;
      const uint32_t off = __shfl_sync(full, excl, (int)(where[u] >> 8));
      if (i < staged) {
        const unsigned long long o = base + off + (where[u] & 0xFFu);
        if (o < a.cap) {
          a.out_key[o] = key[u];
          a.out_dist[o] = st_dist[i];
        }
      }
    }
    __syncwarp();
    staged = 0;
  };

  for (;;) {
    uint32_t q0 = 0;
    if (lane == 0) q0 = atomicAdd(&a.ctr->work, grab);
    q0 = __shfl_sync(full, q0, 0);
    if (q0 >= nq) break;
    const uint32_t q1 = min(q0 + grab, nq);
    for (uint32_t qi = q0; qi < q1; ++qi) {
      // ---- locate the query
      float q[kDim];
      uint64_t qk;
      if (STAGE) {
#pragma unroll
        for (int d = 0; d < kDim; ++d) q[d] = __ldg(a.features + (size_t)qi * kDim + d);
        qk = (uint64_t)qi << 32;
      } else {
        if (qi < e_q0 || qi >= e_q1) {
          flush();
          entry = find_entry(a.q_off, a.B, qi, lane);
          staged_entry = entry;
          e_q0 = __ldg(a.q_off + entry);
          e_q1 = __ldg(a.q_off + entry + 1);
          slot = __ldg(a.entry_slot + entry);
          ev_off = a.slots[slot].num_events;  // query_start_offset
          frow = __ldg(a.feat_row + entry);
        }
        const uint32_t p = (uint32_t)a.step * (qi - e_q0 + 1u);  // seeds at step, 2*step, ... (Q3)
        const float *f = a.features + (size_t)frow * kFeatCap + p;
#pragma unroll
        for (int d = 0; d < kDim; ++d) q[d] = __ldg(f + d);
        qk = a.key.pack(entry, 0u, 0u, p + ev_off);
      }
      __half2 qh01, qh23, qh45;
      float theta16 = 0.0f;
      if (BOX16) {
        qh01 = __floats2half2_rn(q[0], q[1]);
        qh23 = __floats2half2_rn(q[2], q[3]);
        qh45 = __floats2half2_rn(q[4], q[5]);
        float qq = 0.0f;
#pragma unroll
        for (int d = 0; d < kDim; ++d) qq = __fmaf_rn(q[d], q[d], qq);
        const float reach = sqrtf(r2) + 4.8828125e-4f * sqrtf(qq) * 1.001f;  // r + 2^-11 |q|_2
        theta16 = reach * reach * 1.0045f;                                   // (1 + 2^-11)^8 < 1.004
      }
      uint32_t qhits = 0;
      bool capped = false;
      // L = lowest level with nodes waiting (n_levels when none), c = how many wait there
      int L = top_level, c = (int)n_top, nleaf = 0;
      uint32_t waiting = 1u << top_level;  // BFS: levels with nodes on their stack
      if (lane < 16) lcnt[lane] = (BFS && lane == top_level) ? n_top : 0u;
      if (lane < (int)n_top) lstk[top_level * kLevelCap + lane] = (uint32_t)lane;
      __syncwarp();
      for (;;) {
        if (BFS) {
          if (nleaf < 8 && waiting) {
            L = 31 - __clz(waiting);
            if (L > 0 && lcnt[L - 1] > (uint32_t)(kLevelCap - 64)) L = __ffs(waiting) - 1;
            c = (int)lcnt[L];
          } else {
            L = waiting ? 0 : n_levels;
          }
        }
        if (nleaf >= 8 || (L >= n_levels && nleaf > 0)) {
          // ---- leaf step: up to eight leaves, two per 8-lane group, one point each per lane
          const int take = min(nleaf, 8);
          const bool hasA = grp < take, hasB = grp + 4 < take;
          const uint32_t leafA = hasA ? leafq[nleaf - 1 - grp] : 0u;
          const uint32_t leafB = hasB ? leafq[nleaf - 5 - grp] : 0u;
          nleaf -= take;
          const float2 *la = ix.leaf_vals + (size_t)leafA * (3 * kLeaf) + sub;
          const float2 *lb = ix.leaf_vals + (size_t)leafB * (3 * kLeaf) + sub;
          const float2 a01 = __ldg(la), a23 = __ldg(la + kLeaf), a45 = __ldg(la + 2 * kLeaf);
          const float2 b01 = __ldg(lb), b23 = __ldg(lb + kLeaf), b45 = __ldg(lb + 2 * kLeaf);
          const float va[kDim] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
          const float vb[kDim] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y};
          const float d2a = exact_d2(q, va), d2b = exact_d2(q, vb);
          const uint32_t hitA = __ballot_sync(full, hasA && d2a < r2);
          const uint32_t hitB = __ballot_sync(full, hasB && d2b < r2);
          __syncwarp();  // leafq reads are done before a later node step pushes
          bool stop = false;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hit = half ? hitB : hitA;
            if (!hit || stop) continue;
            if (!STAGE) {
              const uint32_t room = kMaxHits - qhits;
              if ((uint32_t)__popc(hit) > room) {
                // keep the first `room` hits in lane order (deterministic)
                uint32_t keep = 0, m = hit;
                for (uint32_t k = 0; k < room; ++k) {
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
              const int at = staged + __popc(hit & lt);
              st_pid[at] = (half ? leafB : leafA) * kLeaf + sub;
              st_dist[at] = half ? d2b : d2a;
              st_qk[at] = qk;
            }
            __syncwarp();
            staged += nh;
            qhits += nh;
            if (!STAGE && qhits >= kMaxHits) {
              capped = true;  // conservatively: there may have been more than 5000
              stop = true;    // the reference stops taking hits after 5000 (spatial_index.cc:371-372)
            }
          }
          if (stop) break;
        } else if (L < n_levels) {
          // ---- node step: up to eight nodes of level L, two per 8-lane group
          const int take = min(c, 8);
          const bool hasA = grp < take, hasB = grp + 4 < take;
          const uint32_t *stk = lstk + L * kLevelCap;
          const uint32_t nodeA = hasA ? stk[c - 1 - grp] : 0u;
          const uint32_t nodeB = hasB ? stk[c - 5 - grp] : 0u;
          c -= take;
          const uint2 *base = ix.level_node[L] + sub;
          const uint2 *ra = base + (size_t)nodeA * (3 * kFan);
          const uint2 *rb = base + (size_t)nodeB * (3 * kFan);
          const uint2 a0 = __ldg(ra), a1 = __ldg(ra + kFan), a2 = __ldg(ra + 2 * kFan);
          const uint2 b0 = __ldg(rb), b1 = __ldg(rb + kFan), b2 = __ldg(rb + 2 * kFan);
          // boxes only prune (stored rounded outwards, tested with slack), so this distance may
          // use FMA; the accept test may not
          float sa = 0.0f, sb = 0.0f;
          if (BOX16) {
            const __half2 zero = __float2half2_rn(0.0f);
            auto as_h2 = [](uint32_t v) { return *reinterpret_cast<const __half2 *>(&v); };
            auto box_d2 = [&](const uint2 r0, const uint2 r1, const uint2 r2_) -> float {
              const __half2 t01 = __hmax2(__hmax2(__hsub2(as_h2(r0.x), qh01), __hsub2(qh01, as_h2(r1.y))), zero);
              const __half2 t23 = __hmax2(__hmax2(__hsub2(as_h2(r0.y), qh23), __hsub2(qh23, as_h2(r2_.x))), zero);
              const __half2 t45 = __hmax2(__hmax2(__hsub2(as_h2(r1.x), qh45), __hsub2(qh45, as_h2(r2_.y))), zero);
              __half2 acc = __hmul2(t01, t01);
              acc = __hfma2(t23, t23, acc);
              acc = __hfma2(t45, t45, acc);
              return __low2float(acc) + __high2float(acc);
            };
            sa = box_d2(a0, a1, a2);
            sb = box_d2(b0, b1, b2);
          } else {
            const float2 l01 = unpack_h2(a0.x), l23 = unpack_h2(a0.y), l45 = unpack_h2(a1.x);
            const float2 h01 = unpack_h2(a1.y), h23 = unpack_h2(a2.x), h45 = unpack_h2(a2.y);
            const float lo[kDim] = {l01.x, l01.y, l23.x, l23.y, l45.x, l45.y};
            const float hi[kDim] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y};
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
              const float t = fmaxf(fmaxf(lo[d] - q[d], q[d] - hi[d]), 0.0f);
              sa = __fmaf_rn(t, t, sa);
            }
          }
          if (!BOX16) {
            const float2 l01 = unpack_h2(b0.x), l23 = unpack_h2(b0.y), l45 = unpack_h2(b1.x);
            const float2 h01 = unpack_h2(b1.y), h23 = unpack_h2(b2.x), h45 = unpack_h2(b2.y);
            const float lo[kDim] = {l01.x, l01.y, l23.x, l23.y, l45.x, l45.y};
            const float hi[kDim] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y};
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
              const float t = fmaxf(fmaxf(lo[d] - q[d], q[d] - hi[d]), 0.0f);
              sb = __fmaf_rn(t, t, sb);
            }
          }
          const float prune_at = BOX16 ? theta16 : r2_prune;
          const uint32_t mA = __ballot_sync(full, hasA && sa <= prune_at);
          const uint32_t mB = __ballot_sync(full, hasB && sb <= prune_at);
          const int nA = __popc(mA), nB = __popc(mB);
          if (BFS) {
            if (L > 0) {
              const uint32_t below = lcnt[L - 1];
              uint32_t *dst = lstk + (L - 1) * kLevelCap + below;
              if (mA & (1u << lane)) dst[__popc(mA & lt)] = nodeA * kFan + sub;
              if (mB & (1u << lane)) dst[nA + __popc(mB & lt)] = nodeB * kFan + sub;
              __syncwarp();  // every lane has read lcnt before lane 0 rewrites it
              if (lane == 0) {
                lcnt[L] = (uint32_t)c;
                lcnt[L - 1] = below + (uint32_t)(nA + nB);
              }
              if (nA + nB) waiting |= 1u << (L - 1);
            } else {
              if (mA & (1u << lane)) leafq[nleaf + __popc(mA & lt)] = nodeA * kFan + sub;
              if (mB & (1u << lane)) leafq[nleaf + nA + __popc(mB & lt)] = nodeB * kFan + sub;
              nleaf += nA + nB;
              __syncwarp();
              if (lane == 0) lcnt[0] = (uint32_t)c;
            }
            if (c == 0) waiting &= ~(1u << L);
            __syncwarp();
            continue;
          }
          if (L > 0) {
            if (nA + nB) {
              // descend: the level below is empty (it is always drained before this one)
              uint32_t *dst = lstk + (L - 1) * kLevelCap;
              if (mA & (1u << lane)) dst[__popc(mA & lt)] = nodeA * kFan + sub;
              if (mB & (1u << lane)) dst[nA + __popc(mB & lt)] = nodeB * kFan + sub;
              lcnt[L] = (uint32_t)c;
              --L;
              c = nA + nB;
            }
          } else {
            if (mA & (1u << lane)) leafq[nleaf + __popc(mA & lt)] = nodeA * kFan + sub;
            if (mB & (1u << lane)) leafq[nleaf + nA + __popc(mB & lt)] = nodeB * kFan + sub;
            nleaf += nA + nB;
          }
          __syncwarp();
          while (c == 0 && ++L < n_levels) c = (int)lcnt[L];  // climb to the next waiting level
        } else {
          break;
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
