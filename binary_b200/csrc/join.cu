// K3 + K4 -- the batched overlap join: per-query bound lookup, candidate scan, count -> prefix sum ->
// scatter of (query_id, target_id) pairs.
//
// Reference being replaced: IntervalTree::find_overlaps / find_overlaps_impl
// (interval_tree.hpp:161-168, 306-334), one recursive pruned walk + vector copies per query, driven
// once per record by sv2nl (mapper.hpp:207-218). Here one thread block takes a tile of 1024
// consecutive queries:
//   1. bounds    lb = dir[bin(q.low)].lb, ub = dir[bin(q.high)+1].ub  (two 8-byte loads; the
//                directory replaces both binary searches, index_build.cu)
//   2. count     rows [lb,ub) are a superset of the hits; the exact predicate
//                q.low <= t.high && t.low <= q.high (interval_tree.hpp:119-121) is evaluated on each.
//                Short ranges: by the owning thread (hit bitmask kept in a register).
//                Long ranges: warp-cooperatively, 32 rows per step, __ballot_sync/__popc.
//   3. prefix    block scan of the counts + decoupled look-back across tiles (lookback.cuh) -> the
//                u64 CSR offsets come out of the SAME kernel, in query order
//   4. scatter   short ranges replay the bitmask; long ranges re-scan with ballot/popc compaction so
//                a warp writes its hits to consecutive addresses
// Modes: kModeFused = 1-4 in one launch; kModeCount = 1-3 (offsets only); kModeScatter = 1,2,4 with
// offsets given (the two-call ABI); kModeAny = 1-2, writes (count > 0).
#include "common.cuh"
#include "lookback.cuh"

namespace bcu {

constexpr int kJoinThreads = 256;
constexpr int kQPT = 4;  // queries per thread: one 128-bit load per input column
constexpr int kJoinTile = kJoinThreads * kQPT;
constexpr uint32_t kScalarMax = 16;  // longer candidate ranges go to the warp-cooperative path

struct JoinArgs {
  const uint2* __restrict__ lowhigh;
  const uint32_t* __restrict__ ids;
  const uint2* __restrict__ dir;
  const GroupDesc* __restrict__ groups;
  uint32_t n_groups;
  uint32_t shift;
  const uint32_t* __restrict__ qgroup;
  const uint32_t* __restrict__ qlow;
  const uint32_t* __restrict__ qhigh;
  uint32_t n_q;
  uint32_t n_tiles;
  int vec_ok;  // all three query columns are 16-byte aligned
  uint64_t* offsets;
  uint64_t capacity;
  uint32_t* __restrict__ hit_query;
  uint32_t* __restrict__ hit_target;
  uint64_t* total;
  uint8_t* __restrict__ any;
  uint32_t qid_base;
  uint64_t* status;
  uint32_t* ticket;
};

__device__ __forceinline__ int find_group(const GroupDesc* groups, uint32_t n_groups, uint32_t g) {
  uint32_t lo = 0, hi = n_groups;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (groups[mid].gval < g) lo = mid + 1; else hi = mid;
  }
  return (lo < n_groups && groups[lo].gval == g) ? (int)lo : -1;
}

__device__ __forceinline__ bool overlaps(uint32_t ql, uint32_t qh, uint2 t) {
  return (ql <= t.y) & (t.x <= qh);
}

template <int MODE>
__global__ void __launch_bounds__(kJoinThreads) join_kernel(const JoinArgs a) {
  __shared__ GroupDesc s_groups[kMaxSmemGroups];
  __shared__ uint64_t s_scan[kJoinThreads / 32 + 1];
  __shared__ uint32_t s_tile;
  __shared__ uint64_t s_base;
  constexpr bool kNeedsPrefix = (MODE == kModeCount || MODE == kModeFused);
  constexpr bool kEmits = (MODE == kModeScatter || MODE == kModeFused);

  const int tid = threadIdx.x, lane = tid & 31;
  if (kNeedsPrefix) {
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
  }
  const bool groups_in_smem = a.n_groups <= (uint32_t)kMaxSmemGroups;
  if (groups_in_smem)
    for (uint32_t g = tid; g < a.n_groups; g += kJoinThreads) s_groups[g] = a.groups[g];
  __syncthreads();
  const uint32_t tile = kNeedsPrefix ? s_tile : blockIdx.x;
  const GroupDesc* groups = groups_in_smem ? s_groups : a.groups;

  // ---- load 4 consecutive queries per thread -------------------------------------------------------
  const uint32_t q0 = tile * (uint32_t)kJoinTile + (uint32_t)tid * kQPT;
  uint32_t ql[kQPT], qh[kQPT], qg[kQPT];
  if (a.vec_ok && q0 + kQPT <= a.n_q) {
    uint4 t = *reinterpret_cast<const uint4*>(a.qlow + q0);
    ql[0] = t.x; ql[1] = t.y; ql[2] = t.z; ql[3] = t.w;
    t = *reinterpret_cast<const uint4*>(a.qhigh + q0);
    qh[0] = t.x; qh[1] = t.y; qh[2] = t.z; qh[3] = t.w;
    if (a.qgroup) {
      t = *reinterpret_cast<const uint4*>(a.qgroup + q0);
      qg[0] = t.x; qg[1] = t.y; qg[2] = t.z; qg[3] = t.w;
    } else {
      qg[0] = qg[1] = qg[2] = qg[3] = 0u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      bool v = q0 + j < a.n_q;
      ql[j] = v ? a.qlow[q0 + j] : 0u;
      qh[j] = v ? a.qhigh[q0 + j] : 0u;
      qg[j] = (v && a.qgroup) ? a.qgroup[q0 + j] : 0u;
    }
  }

  // ---- 1. bounds -------------------------------------------------------------------------------------
  uint32_t lb[kQPT], ub[kQPT];
#pragma unroll
  for (int j = 0; j < kQPT; ++j) {
    lb[j] = ub[j] = 0;
    if (q0 + j < a.n_q) {
      int gi = find_group(groups, a.n_groups, qg[j]);
      if (gi >= 0) {
        const uint32_t nb = groups[gi].nb;
        const uint64_t bin_base = groups[gi].bin_base;
        const uint32_t b_lo = ql[j] >> a.shift;
        if (b_lo < nb) {  // otherwise q.low lies beyond every high of the group
          uint32_t b_hi1 = (qh[j] >> a.shift) + 1u;
          if (b_hi1 > nb) b_hi1 = nb;
          lb[j] = a.dir[bin_base + b_lo].x;
          ub[j] = a.dir[bin_base + b_hi1].y;
          if (ub[j] < lb[j]) ub[j] = lb[j];
        }
      }
    }
  }

  // ---- 2. count ----------------------------------------------------------------------------------------
  uint32_t cnt[kQPT], mask[kQPT];
#pragma unroll
  for (int j = 0; j < kQPT; ++j) {
    cnt[j] = 0;
    mask[j] = 0;
    const uint32_t len = ub[j] - lb[j];
    if (len <= kScalarMax) {
      for (uint32_t k = 0; k < len; ++k) {
        uint2 t = a.lowhigh[lb[j] + k];
        mask[j] |= (uint32_t)overlaps(ql[j], qh[j], t) << k;
      }
      cnt[j] = __popc(mask[j]);
    }
  }
  if (MODE != kModeScatter) {
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      unsigned big = __ballot_sync(0xffffffffu, ub[j] - lb[j] > kScalarMax);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t blb = __shfl_sync(0xffffffffu, lb[j], src);
        const uint32_t bub = __shfl_sync(0xffffffffu, ub[j], src);
        const uint32_t bql = __shfl_sync(0xffffffffu, ql[j], src);
        const uint32_t bqh = __shfl_sync(0xffffffffu, qh[j], src);
        uint32_t c = 0;
        for (uint32_t r = blb + lane; r < bub; r += 32) c += overlaps(bql, bqh, a.lowhigh[r]);
#pragma unroll
        for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (lane == src) cnt[j] = c;
      }
    }
  }

  if (MODE == kModeAny) {
#pragma unroll
    for (int j = 0; j < kQPT; ++j)
      if (q0 + j < a.n_q) a.any[q0 + j] = cnt[j] ? 1 : 0;
    return;
  }

  // ---- 3. prefix sum -> CSR offsets --------------------------------------------------------------------
  uint64_t off[kQPT];
  if (kNeedsPrefix) {
    const uint64_t thread_sum = (uint64_t)cnt[0] + cnt[1] + cnt[2] + cnt[3];
    uint64_t block_total;
    uint64_t excl = block_exclusive_scan<SumOp, kJoinThreads>(thread_sum, s_scan, &block_total);
    if (tid < 32) {
      uint64_t e = lookback_exclusive<SumOp>(a.status, tile, block_total);
      if (tid == 0) {
        s_base = e;
        if (tile == a.n_tiles - 1) {
          a.offsets[a.n_q] = e + block_total;
          if (a.total) *a.total = e + block_total;
        }
      }
    }
    __syncthreads();
    uint64_t run = s_base + excl;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) { off[j] = run; run += cnt[j]; }
    if (a.vec_ok && q0 + kQPT <= a.n_q) {
      ulonglong2* o = reinterpret_cast<ulonglong2*>(a.offsets + q0);
      o[0] = make_ulonglong2(off[0], off[1]);
      o[1] = make_ulonglong2(off[2], off[3]);
    } else {
#pragma unroll
      for (int j = 0; j < kQPT; ++j) if (q0 + j < a.n_q) a.offsets[q0 + j] = off[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < kQPT; ++j) off[j] = (q0 + j < a.n_q) ? a.offsets[q0 + j] : 0ull;
  }
  if (!kEmits) return;

  // ---- 4. scatter ----------------------------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < kQPT; ++j) {
    uint32_t m = mask[j];
    uint64_t pos = off[j];
    const uint32_t qid = a.qid_base + q0 + j;
    while (m) {
      const uint32_t k = __ffs(m) - 1;
      m &= m - 1;
      if (pos < a.capacity) {
        a.hit_target[pos] = a.ids[lb[j] + k];
        a.hit_query[pos] = qid;
      }
      ++pos;
    }
  }
#pragma unroll
  for (int j = 0; j < kQPT; ++j) {
    unsigned big = __ballot_sync(0xffffffffu, ub[j] - lb[j] > kScalarMax);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t blb = __shfl_sync(0xffffffffu, lb[j], src);
      const uint32_t bub = __shfl_sync(0xffffffffu, ub[j], src);
      const uint32_t bql = __shfl_sync(0xffffffffu, ql[j], src);
      const uint32_t bqh = __shfl_sync(0xffffffffu, qh[j], src);
      uint64_t base = shfl_u64(off[j], src);
      const uint32_t qid = a.qid_base + (q0 - (uint32_t)lane * kQPT) + (uint32_t)src * kQPT + j;
      for (uint32_t r0 = blb; r0 < bub; r0 += 32) {
        const uint32_t r = r0 + lane;
        const bool hit = (r < bub) && overlaps(bql, bqh, a.lowhigh[r < bub ? r : blb]);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const uint64_t pos = base + __popc(bal & ((1u << lane) - 1u));
          if (pos < a.capacity) {
            a.hit_target[pos] = a.ids[r];
            a.hit_query[pos] = qid;
          }
        }
        base += __popc(bal);
      }
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_join(const bcu_index* ix, int mode, uint64_t n_q, const uint32_t* d_qgroup,
                const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                uint64_t pair_capacity, uint32_t* d_hit_query, uint32_t* d_hit_target,
                uint64_t* d_total, uint8_t* d_any, uint32_t query_id_base, cudaStream_t stream) {
  if (n_q > 0xfffffffeull) { set_error("query batch exceeds 2^32-2 queries"); return BCU_E_LIMIT; }
  const bool prefix = (mode == kModeCount || mode == kModeFused);
  if (n_q == 0 || ix->n == 0) {  // nothing can hit: offsets are all zero
    if (prefix) {
      BCU_CUDA(cudaMemsetAsync(d_offsets, 0, (n_q + 1) * 8, stream));
      if (d_total) BCU_CUDA(cudaMemsetAsync(d_total, 0, 8, stream));
    }
    if (mode == kModeAny && n_q) BCU_CUDA(cudaMemsetAsync(d_any, 0, n_q, stream));
    return BCU_OK;
  }
  JoinArgs a;
  a.lowhigh = ix->d_lowhigh;
  a.ids = ix->d_id;
  a.dir = ix->d_dir;
  a.groups = ix->d_groups;
  a.n_groups = ix->n_groups;
  a.shift = ix->shift;
  a.qgroup = d_qgroup;
  a.qlow = d_qlow;
  a.qhigh = d_qhigh;
  a.n_q = (uint32_t)n_q;
  a.n_tiles = (uint32_t)((n_q + kJoinTile - 1) / kJoinTile);
  a.vec_ok = aligned16(d_qlow) && aligned16(d_qhigh) && (!d_qgroup || aligned16(d_qgroup)) &&
             (!d_offsets || aligned16(d_offsets));
  a.offsets = d_offsets;
  a.capacity = (mode == kModeScatter) ? ~0ull : pair_capacity;
  a.hit_query = d_hit_query;
  a.hit_target = d_hit_target;
  a.total = d_total;
  a.any = d_any;
  a.qid_base = query_id_base;
  a.status = nullptr;
  a.ticket = nullptr;
  void* scratch = nullptr;
  if (prefix) {
    size_t bytes = ((size_t)a.n_tiles + 1) * 8;
    BCU_CUDA(cudaMallocAsync(&scratch, bytes, stream));
    BCU_CUDA(cudaMemsetAsync(scratch, 0, bytes, stream));
    a.status = reinterpret_cast<uint64_t*>(scratch);
    a.ticket = reinterpret_cast<uint32_t*>(a.status + a.n_tiles);
  }
  switch (mode) {
    case kModeCount: join_kernel<kModeCount><<<a.n_tiles, kJoinThreads, 0, stream>>>(a); break;
    case kModeScatter: join_kernel<kModeScatter><<<a.n_tiles, kJoinThreads, 0, stream>>>(a); break;
    case kModeFused: join_kernel<kModeFused><<<a.n_tiles, kJoinThreads, 0, stream>>>(a); break;
    case kModeAny: join_kernel<kModeAny><<<a.n_tiles, kJoinThreads, 0, stream>>>(a); break;
    default: set_error("bad join mode %d", mode); return BCU_E_INVALID;
  }
  BCU_LAUNCHED();
  if (scratch) BCU_CUDA(cudaFreeAsync(scratch, stream));
  return BCU_OK;
}

}  // namespace bcu
