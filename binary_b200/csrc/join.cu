// K3 + K4 -- the batched overlap join: per-query bound lookup, candidate scan, count -> prefix sum ->
// scatter of (query_id, target_id) pairs.
//
// Reference being replaced: IntervalTree::find_overlaps / find_overlaps_impl
// (interval_tree.hpp:161-168, 306-334), one recursive pruned walk + vector copies per query, driven
// once per record by sv2nl (mapper.hpp:207-218).
//
// Execution model: a persistent grid (one wave of resident CTAs); every WARP owns a tile of 128
// consecutive queries (4 per lane, one 128-bit load per input column) and loops over tiles
// warp-stride. There is no block-level synchronisation in the loop: warps run fully decoupled, which
// is what hides the three dependent L2 round trips of a query (directory -> candidate rows -> ids).
//   1. bounds    lb = dir[bin(q.low)].lb, ub = dir[bin(q.high)+1].ub  (two 8-byte loads; the
//                directory replaces both binary searches, index_build.cu)
//   2. count     rows [lb,ub) are a superset of the hits; the exact predicate
//                q.low <= t.high && t.low <= q.high (interval_tree.hpp:119-121) is evaluated on each.
//                Short ranges: by the owning lane, the 4 queries of a lane interleaved so their loads
//                overlap; the hit bitmask stays in a register.
//                Long ranges: warp-cooperatively, 32 rows per step, __ballot_sync/__popc.
//   3. prefix    warp scan of the counts + decoupled look-back across tiles (lookback.cuh): the u64 CSR
//                offsets come out of the SAME kernel, in query order
//   4. scatter   short ranges: hit rows are staged in shared memory at their tile-local rank, then the
//                warp writes (query_id, target_id) to consecutive addresses; long ranges re-scan with
//                ballot/popc compaction.
// Modes: kModeFused = 1-4 in one launch; kModeCount = 1-3 (offsets only); kModeScatter = 1,2,4 with
// offsets given (the two-call ABI); kModeAny = 1-2, writes (count > 0).
#include "common.cuh"
#include "lookback.cuh"

namespace bcu {

constexpr int kJoinThreads = 256;
constexpr int kJoinWarps = kJoinThreads / 32;
constexpr int kQPT = 4;                  // queries per lane: one 128-bit load per input column
constexpr int kWarpTile = 32 * kQPT;     // 128 queries per warp tile
constexpr uint32_t kScalarMax = 16;      // longer candidate ranges go to the warp-cooperative path
constexpr int kStage = 256;              // staged hits per warp and round
constexpr int kDirectGroups = 1024;      // group values below this use a direct map

struct JoinArgs {
  const uint2* __restrict__ lowhigh;
  const uint32_t* __restrict__ ids;
  const uint2* __restrict__ dir;
  const GroupDesc* __restrict__ groups;
  uint32_t n_groups;
  uint32_t shift;
  uint32_t max_gval;
  const uint32_t* __restrict__ qgroup;
  const uint32_t* __restrict__ qlow;
  const uint32_t* __restrict__ qhigh;
  uint32_t n_q;
  uint32_t n_tiles;
  int vec_ok;  // all query columns (and offsets) are 16-byte aligned
  uint64_t* offsets;
  uint64_t capacity;
  uint32_t* __restrict__ hit_query;
  uint32_t* __restrict__ hit_target;
  uint64_t* total;
  uint8_t* __restrict__ any;
  uint32_t qid_base;
  uint64_t* status;
};

struct JoinSmem {
  uint32_t g_val[kMaxSmemGroups];   // sorted group values (binary-search mode)
  uint32_t g_nb[kMaxSmemGroups];
  uint64_t g_base[kMaxSmemGroups];
  uint16_t g_map[kDirectGroups];    // group value -> descriptor index, 0xffff = absent (direct mode)
  uint64_t st_pos[kJoinWarps][kStage];
  uint32_t st_row[kJoinWarps][kStage];
  uint8_t st_qi[kJoinWarps][kStage];
};

__device__ __forceinline__ bool overlaps(uint32_t ql, uint32_t qh, uint2 t) {
  return (ql <= t.y) & (t.x <= qh);
}

__device__ __forceinline__ void load_queries(const JoinArgs& a, uint32_t q0, uint32_t (&ql)[kQPT],
                                             uint32_t (&qh)[kQPT], uint32_t (&qg)[kQPT]) {
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
      bool v = q0 + j < a.n_q;  // q0 may be past the end for the prefetch of a non-existent tile
      ql[j] = v ? a.qlow[q0 + j] : 0u;
      qh[j] = v ? a.qhigh[q0 + j] : 0u;
      qg[j] = (v && a.qgroup) ? a.qgroup[q0 + j] : 0u;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kJoinThreads) join_kernel(const JoinArgs a) {
  __shared__ JoinSmem sm;
  constexpr bool kNeedsPrefix = (MODE == kModeCount || MODE == kModeFused);
  constexpr bool kEmits = (MODE == kModeScatter || MODE == kModeFused);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- group tables -> shared memory, once per CTA --------------------------------------------------
  const bool direct = a.max_gval < (uint32_t)kDirectGroups && a.n_groups <= (uint32_t)kMaxSmemGroups;
  const bool in_smem = a.n_groups <= (uint32_t)kMaxSmemGroups;
  if (in_smem) {
    if (direct)
      for (int g = tid; g < kDirectGroups; g += kJoinThreads) sm.g_map[g] = 0xffffu;
    __syncthreads();
    for (uint32_t g = tid; g < a.n_groups; g += kJoinThreads) {
      const GroupDesc d = a.groups[g];
      sm.g_val[g] = d.gval;
      sm.g_nb[g] = d.nb;
      sm.g_base[g] = d.bin_base;
      if (direct) sm.g_map[d.gval] = (uint16_t)g;
    }
  }
  __syncthreads();

  const uint32_t warps_total = gridDim.x * kJoinWarps;
  uint32_t tile = blockIdx.x * kJoinWarps + warp;

  uint32_t nql[kQPT], nqh[kQPT], nqg[kQPT];  // prefetched queries of the next tile
  if (tile < a.n_tiles) load_queries(a, tile * kWarpTile + lane * kQPT, nql, nqh, nqg);

  for (; tile < a.n_tiles; tile += warps_total) {
    const uint32_t q0 = tile * (uint32_t)kWarpTile + (uint32_t)lane * kQPT;
    uint32_t ql[kQPT], qh[kQPT], qg[kQPT];
#pragma unroll
    for (int j = 0; j < kQPT; ++j) { ql[j] = nql[j]; qh[j] = nqh[j]; qg[j] = nqg[j]; }
    if (tile + warps_total < a.n_tiles)
      load_queries(a, (tile + warps_total) * kWarpTile + lane * kQPT, nql, nqh, nqg);

    // ---- 1. bounds ---------------------------------------------------------------------------------
    uint32_t lb[kQPT], len[kQPT];
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      lb[j] = 0;
      len[j] = 0;
      uint32_t nb = 0;
      uint64_t bin_base = 0;
      if (q0 + j < a.n_q) {
        if (direct) {
          const uint32_t gi = qg[j] < (uint32_t)kDirectGroups ? sm.g_map[qg[j]] : 0xffffu;
          if (gi != 0xffffu) { nb = sm.g_nb[gi]; bin_base = sm.g_base[gi]; }
        } else if (in_smem) {
          uint32_t lo = 0, hi = a.n_groups;
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (sm.g_val[mid] < qg[j]) lo = mid + 1; else hi = mid;
          }
          if (lo < a.n_groups && sm.g_val[lo] == qg[j]) { nb = sm.g_nb[lo]; bin_base = sm.g_base[lo]; }
        } else {
          uint32_t lo = 0, hi = a.n_groups;
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (a.groups[mid].gval < qg[j]) lo = mid + 1; else hi = mid;
          }
          if (lo < a.n_groups && a.groups[lo].gval == qg[j]) { nb = a.groups[lo].nb; bin_base = a.groups[lo].bin_base; }
        }
      }
      const uint32_t b_lo = ql[j] >> a.shift;
      if (b_lo < nb) {  // otherwise unknown group, or q.low lies beyond every high of the group
        uint32_t b_hi1 = (qh[j] >> a.shift) + 1u;
        if (b_hi1 > nb) b_hi1 = nb;
        const uint32_t l = a.dir[bin_base + b_lo].x;
        const uint32_t u = a.dir[bin_base + b_hi1].y;
        lb[j] = l;
        len[j] = u > l ? u - l : 0u;
      }
    }

    // ---- 2. count ----------------------------------------------------------------------------------
    uint32_t cnt[kQPT], mask[kQPT];
    uint32_t max_short = 0;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      cnt[j] = 0;
      mask[j] = 0;
      if (len[j] <= kScalarMax && len[j] > max_short) max_short = len[j];
    }
    // the 4 queries of a lane advance together: 4 independent loads per trip
    for (uint32_t k = 0; k < max_short; ++k) {
#pragma unroll
      for (int j = 0; j < kQPT; ++j) {
        if (k < len[j] && len[j] <= kScalarMax) {
          const uint2 t = a.lowhigh[lb[j] + k];
          mask[j] |= (uint32_t)overlaps(ql[j], qh[j], t) << k;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kQPT; ++j) cnt[j] = __popc(mask[j]);
    bool any_big = false;
    if (MODE != kModeScatter) {
#pragma unroll
      for (int j = 0; j < kQPT; ++j) {
        unsigned big = __ballot_sync(0xffffffffu, len[j] > kScalarMax);
        any_big |= (big != 0);
        while (big) {
          const int src = __ffs(big) - 1;
          big &= big - 1;
          const uint32_t blb = __shfl_sync(0xffffffffu, lb[j], src);
          const uint32_t bub = blb + __shfl_sync(0xffffffffu, len[j], src);
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
      continue;
    }

    // ---- 3. prefix sum -> CSR offsets --------------------------------------------------------------
    uint64_t off[kQPT];
    if (kNeedsPrefix) {
      const uint64_t lane_sum = (uint64_t)cnt[0] + cnt[1] + cnt[2] + cnt[3];
      uint64_t incl = lane_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t v = shfl_up_u64(incl, o);
        if (lane >= o) incl += v;
      }
      const uint64_t tile_total = shfl_u64(incl, 31);
      const uint64_t base = lookback_exclusive<SumOp>(a.status, tile, tile_total);
      if (tile == a.n_tiles - 1 && lane == 0) {
        a.offsets[a.n_q] = base + tile_total;
        if (a.total) *a.total = base + tile_total;
      }
      uint64_t run = base + incl - lane_sum;
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
    if (!kEmits) continue;

    // ---- 4. scatter --------------------------------------------------------------------------------
    // short ranges: rank of each hit among the tile's short-range hits -> staging slot
    {
      const uint32_t lane_hits = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]);
      uint32_t incl = lane_hits;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t staged_total = __shfl_sync(0xffffffffu, incl, 31);
      const uint32_t first_slot = incl - lane_hits;
      const uint32_t qid0 = a.qid_base + tile * (uint32_t)kWarpTile;
      for (uint32_t r0 = 0; r0 < staged_total; r0 += kStage) {
        uint32_t slot = first_slot;
#pragma unroll
        for (int j = 0; j < kQPT; ++j) {
          uint32_t m = mask[j];
          uint64_t pos = off[j];
          while (m) {
            const uint32_t k = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t s = slot - r0;  // wraps below the window: caught by the unsigned compare
            if (s < (uint32_t)kStage) {
              sm.st_row[warp][s] = lb[j] + k;
              sm.st_pos[warp][s] = pos;
              sm.st_qi[warp][s] = (uint8_t)(lane * kQPT + j);
            }
            ++slot;
            ++pos;
          }
        }
        __syncwarp();
        const uint32_t n_here = min((uint32_t)kStage, staged_total - r0);
        for (uint32_t s = lane; s < n_here; s += 32) {
          const uint64_t pos = sm.st_pos[warp][s];
          if (pos < a.capacity) {
            a.hit_target[pos] = a.ids[sm.st_row[warp][s]];
            a.hit_query[pos] = qid0 + sm.st_qi[warp][s];
          }
        }
        __syncwarp();
      }
    }
    // long ranges: warp-cooperative re-scan with ballot/popc compaction
    if (MODE == kModeScatter) {
#pragma unroll
      for (int j = 0; j < kQPT; ++j) any_big |= __any_sync(0xffffffffu, len[j] > kScalarMax);
    }
    if (any_big) {
#pragma unroll
      for (int j = 0; j < kQPT; ++j) {
        unsigned big = __ballot_sync(0xffffffffu, len[j] > kScalarMax);
        while (big) {
          const int src = __ffs(big) - 1;
          big &= big - 1;
          const uint32_t blb = __shfl_sync(0xffffffffu, lb[j], src);
          const uint32_t bub = blb + __shfl_sync(0xffffffffu, len[j], src);
          const uint32_t bql = __shfl_sync(0xffffffffu, ql[j], src);
          const uint32_t bqh = __shfl_sync(0xffffffffu, qh[j], src);
          uint64_t base = shfl_u64(off[j], src);
          const uint32_t qid = a.qid_base + tile * (uint32_t)kWarpTile + (uint32_t)src * kQPT + j;
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
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// resident CTAs per SM of each mode x SM count, per device (queried once)
static int persistent_grid(int device, int mode) {
  static std::atomic<int> cache[64][4];
  if (device >= 0 && device < 64) {
    int v = cache[device][mode].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int per_sm = 0, sms = 0;
  cudaError_t e = cudaSuccess;
  switch (mode) {
    case kModeCount: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, join_kernel<kModeCount>, kJoinThreads, 0); break;
    case kModeScatter: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, join_kernel<kModeScatter>, kJoinThreads, 0); break;
    case kModeFused: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, join_kernel<kModeFused>, kJoinThreads, 0); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, join_kernel<kModeAny>, kJoinThreads, 0); break;
  }
  if (e != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int v = per_sm * sms;
  if (device >= 0 && device < 64 && v > 0) cache[device][mode].store(v, std::memory_order_relaxed);
  return v;
}

int launch_join(const bcu_index* ix, int mode, uint64_t n_q, const uint32_t* d_qgroup,
                const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                uint64_t pair_capacity, uint32_t* d_hit_query, uint32_t* d_hit_target,
                uint64_t* d_total, uint8_t* d_any, uint32_t query_id_base, cudaStream_t stream) {
  if (n_q > 0xfffffffeull) { set_error("query batch exceeds 2^32-2 queries"); return BCU_E_LIMIT; }
  if (mode < 0 || mode > 3) { set_error("bad join mode %d", mode); return BCU_E_INVALID; }
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
  a.max_gval = ix->max_gval;
  a.qgroup = d_qgroup;
  a.qlow = d_qlow;
  a.qhigh = d_qhigh;
  a.n_q = (uint32_t)n_q;
  a.n_tiles = (uint32_t)((n_q + kWarpTile - 1) / kWarpTile);
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
  // Persistent launch: the look-back spins on predecessor tiles, so every CTA must be resident.
  const int resident = persistent_grid(ix->device, mode);
  if (resident <= 0) { set_error("cannot determine a resident grid for the join kernel"); return BCU_E_CUDA; }
  const uint32_t ctas_needed = (a.n_tiles + kJoinWarps - 1) / kJoinWarps;
  const uint32_t grid = ctas_needed < (uint32_t)resident ? ctas_needed : (uint32_t)resident;
  void* scratch = nullptr;
  if (prefix) {
    size_t bytes = (size_t)a.n_tiles * 8;
    BCU_CUDA(cudaMallocAsync(&scratch, bytes, stream));
    BCU_CUDA(cudaMemsetAsync(scratch, 0, bytes, stream));
    a.status = reinterpret_cast<uint64_t*>(scratch);
  }
  switch (mode) {
    case kModeCount: join_kernel<kModeCount><<<grid, kJoinThreads, 0, stream>>>(a); break;
    case kModeScatter: join_kernel<kModeScatter><<<grid, kJoinThreads, 0, stream>>>(a); break;
    case kModeFused: join_kernel<kModeFused><<<grid, kJoinThreads, 0, stream>>>(a); break;
    default: join_kernel<kModeAny><<<grid, kJoinThreads, 0, stream>>>(a); break;
  }
  BCU_LAUNCHED();
  if (scratch) BCU_CUDA(cudaFreeAsync(scratch, stream));
  return BCU_OK;
}

}  // namespace bcu
