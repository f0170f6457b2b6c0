// The BINNED join: the same count -> prefix-sum -> scatter join as join.cu, for batches so large and indexes
// so far beyond L2 (BASELINE config D: 10 M targets, 100 M unsorted queries) that the per-query gathers of the
// general path are what limits it (ncu, round 1: 311 B of DRAM sectors per query, 5.2x the algorithmic bytes).
// Here the QUERIES are brought to the index instead:
//
//   bin_sort_kernel      the batch is cut into tiles of 4096 queries; every tile is counting-sorted IN PLACE by
//                        the index bin (common.cuh BinDesc: a coordinate range of one group whose rows fit a
//                        CTA's shared memory) of its queries. Reads the three query columns once, writes
//                        {low, high} + the query's position in its tile; fully coalesced both ways.
//   bin_transpose_kernel [tile][bin] run starts -> [bin][tile] run descriptors.
//   bin_probe_kernel     persistent CTAs take bins from a ticket. A bin's rows (low / high / id columns of every
//                        length class) are moved into shared memory with cp.async.bulk (1-D TMA copies completed
//                        on an mbarrier), a sub-cell table per class is built, and all the runs of that bin --
//                        one per tile -- are answered from shared memory: per class two table look-ups bound the
//                        candidate window [first row with low >= q.low - maxlen, first row with low > q.high),
//                        every row in it is tested with the exact predicate q.low <= t.high (interval_tree.hpp:
//                        119-121; t.low <= q.high holds by construction), hits are compacted per warp and
//                        written as target ids to a staging area in run order. Queries whose window exceeds the
//                        32-row hit mask or that reach past their bin are listed for bin_spill_kernel.
//   bin_spill_kernel     those few queries, one warp each, through the general index (directory + row scan).
//   bin_place_kernel     per tile: hit counts back into query order (shared memory), prefix sum + decoupled
//                        look-back over the tiles -> u64 CSR offsets; the tile's target ids are gathered run by
//                        run from the staging area into their final order in shared memory and written out
//                        coalesced, then the query-id column the same way.
//
// Results are the general path's: offsets, total, pairs sorted by query id (order inside a query differs, as the
// ABI allows). Traffic per query of config D: ~20 B routing + 8 B probe input + 8 B state (+ 4 B per hit staged
// and read back) on top of the algorithmic bytes, all of it streamed; nothing is gathered from DRAM per query.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "lookback.cuh"

namespace bcu {

constexpr int kTileQ = 4096;          // queries per routing tile
constexpr int kSortThreads = 512;
constexpr int kSortQPT = kTileQ / kSortThreads;
constexpr int kProbeThreads = 1024;
constexpr int kProbeWarps = kProbeThreads / 32;
constexpr int kStageIds = 256;        // hits staged per warp and round
constexpr uint32_t kSlabIds = 8192;   // staging is reserved per warp in slabs: one global atomic per ~40 rounds
constexpr int kPlaceThreads = 512;
constexpr int kPlaceQPT = kTileQ / kPlaceThreads;
#ifndef BCU_PLACE_LOADS
#define BCU_PLACE_LOADS 12
#endif
// ids a lane has in flight in the place kernel's gather: one trip covers nearly every list of the north-star
// workload (99th percentile 14 hits). Measured per step: 4 -> 8.91 ms, 8 -> 8.95, 12 -> 8.79, 16 -> 8.81.
constexpr uint32_t kPlaceLoads = BCU_PLACE_LOADS;
constexpr int kPlaceCap = 16384;      // target ids assembled per tile and round
constexpr uint32_t kMaskRows = 32;    // candidate window a lane can record (one hit-mask word per class)
constexpr int kBnDirectGroups = 1024; // group values below this are routed through a direct map
constexpr uint32_t kNotStored = 0xffffffffu;  // sres.x of a hit list that did not fit the staging area (the join
                                              // exceeds the caller's pair capacity; staging indices stay below it)

struct BinnedArgs {
  // index
  const uint32_t* __restrict__ low;   // (group, low)-sorted columns the bins' own rows are copied from
  const uint32_t* __restrict__ high;
  const uint32_t* __restrict__ ids;
  const unsigned char* __restrict__ blob;  // per bin: sub_start | cov_rel | cov_high | cov_id
  const uint2* __restrict__ lowhigh;  // general index (spill path)
  const uint32_t* __restrict__ gids;
  const DirEntry* __restrict__ dir;
  const GroupDesc* __restrict__ gtable;  // [n_cls][n_groups] (general index: used by the spill path)
  const BinDesc* __restrict__ desc;
  const BinGroup* __restrict__ groups;
  const uint32_t* __restrict__ cellbits;  // [2 * ceil(n_cells / 32)]: first-cell-of-a-bin bits, then their rank per word
  uint32_t n_groups, n_bins, n_cells, cell_shift, max_gval, n_cls;
  // batch
  const uint32_t* __restrict__ qgroup;
  const uint32_t* __restrict__ qlow;
  const uint32_t* __restrict__ qhigh;
  uint32_t n_q, n_tiles;
  // scratch
  uint2* brec;       // [n_tiles * kTileQ] {low, high}, tile by tile, sorted by bin inside a tile
  uint16_t* bloc;    // [n_tiles * kTileQ] position of the slot's query inside its tile
  uint16_t* trun;    // [n_tiles][n_bins + 2] first slot of bin k; [n_bins] = first slot without a bin; [n_bins+1] = queries
  uint32_t* brun;    // [n_bins][n_tiles] first slot | run length << 16
  uint2* sres;       // [n_tiles * kTileQ] {first staged id, hit count}
  uint32_t* staging;
  uint64_t stage_cap;
  unsigned long long* stage_cursor;
  uint32_t* probe_ticket;
  uint32_t* spill;   // [2 * n_q] {slot, bin}
  uint32_t* n_spill;
  uint64_t* status;  // [n_tiles] look-back words of the place kernel
  uint32_t* place_ticket;
  // outputs
  uint64_t* offsets;
  uint64_t capacity;
  uint32_t* hit_query;
  uint32_t* hit_target;
  uint64_t* total;
  uint64_t* total_mapped;
  uint32_t qid_base;
  int emit;          // 0 = offsets only (count mode)
  int ht_aligned, hq_aligned;  // the pair columns are 16-byte aligned (vector stores)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += y;
  }
  return v;
}
// Owner of flattened element f among 32 runs with inclusive prefix `incl` (one per lane): the number of lanes
// whose inclusive prefix is <= f. Called by all lanes; meaningful when f < incl of lane 31.
__device__ __forceinline__ int run_owner(uint32_t incl, uint32_t f) {
  int pos = 0;
#pragma unroll
  for (int step = 16; step; step >>= 1) {
    const uint32_t v = __shfl_sync(0xffffffffu, incl, pos + step - 1);
    if (v <= f) pos += step;
  }
  return pos & 31;
}

// ---------------------------------------------------------------------------------------------------
// Routing: tile-local counting sort by bin. Persistent CTAs (the routing tables are staged once per CTA).
// Ranks come from per-warp histograms updated WITHOUT atomics: the lanes of a warp that hold the same bin find
// each other with match.any, the first of them adds the group's size to the warp's
// counter (shared-memory atomics cost ~2 cycles per LANE on this machine: 4096 of them per tile were 80 % of
// this kernel's time). With more than kSortMaxWarpHistBins bins the per-warp histograms do not fit and
// shared-memory atomics are used instead.
// Dynamic shared memory: rec[kTileQ] uint2 | hist[K + 2] u32 | groups[n_groups] BinGroup | loc[kTileQ] u16 |
// cellbits[2 * ceil(n_cells / 32)] u32 | gmap[kBnDirectGroups] u16 | whist[kSortWarps][K + 2] u16 (WARP_HIST only)
constexpr int kSortWarps = kSortThreads / 32;
constexpr uint32_t kSortMaxWarpHistBins = 2046;

template <bool WARP_HIST>
__global__ void __launch_bounds__(kSortThreads, 2) bin_sort_kernel(const BinnedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t K = a.n_bins;
  const uint32_t Kp = (K + 2 + 3) & ~3u;
  uint2* s_rec = reinterpret_cast<uint2*>(smem_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_rec + kTileQ);
  BinGroup* s_grp = reinterpret_cast<BinGroup*>(s_hist + Kp);
  uint16_t* s_loc = reinterpret_cast<uint16_t*>(s_grp + a.n_groups);
  const uint32_t n_words = (a.n_cells + 31) / 32;
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_loc + kTileQ);  // [n_words] bits, [n_words] ranks
  uint16_t* s_gmap = reinterpret_cast<uint16_t*>(s_bits + ((2 * n_words + 3) & ~3u));
  uint16_t* s_whist = s_gmap + kBnDirectGroups;  // [kSortWarps][Kp]
  __shared__ uint64_t s_scan[kSortThreads / 32 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool direct = a.max_gval < (uint32_t)kBnDirectGroups;

  for (uint32_t i = tid; i < 2 * n_words; i += kSortThreads) s_bits[i] = a.cellbits[i];
  for (uint32_t i = tid; i < a.n_groups; i += kSortThreads) s_grp[i] = a.groups[i];
  if (direct)
    for (int i = tid; i < kBnDirectGroups; i += kSortThreads) s_gmap[i] = 0xffffu;
  __syncthreads();
  if (direct)
    for (uint32_t i = tid; i < a.n_groups; i += kSortThreads) s_gmap[s_grp[i].gval] = (uint16_t)i;

  for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    __syncthreads();  // previous tile's copy-out is done; the group map is visible
    if (WARP_HIST) {
      uint32_t* w32 = reinterpret_cast<uint32_t*>(s_whist);
      for (uint32_t i = tid; i < kSortWarps * Kp / 2; i += kSortThreads) w32[i] = 0;
    } else {
      for (uint32_t i = tid; i < K + 2; i += kSortThreads) s_hist[i] = 0;
    }
    __syncthreads();
    const uint64_t q0 = (uint64_t)tile * kTileQ;
    uint32_t ql[kSortQPT], qh[kSortQPT], qg[kSortQPT], bin[kSortQPT], rank[kSortQPT];
#pragma unroll
    for (int j = 0; j < kSortQPT; ++j) {  // all the loads of the tile first
      const uint64_t q = q0 + (uint32_t)(j * kSortThreads + tid);
      ql[j] = qh[j] = qg[j] = 0;
      if (q < a.n_q) {
        ql[j] = a.qlow[q];
        qh[j] = a.qhigh[q];
        if (a.qgroup) qg[j] = a.qgroup[q];
      }
    }
#pragma unroll
    for (int j = 0; j < kSortQPT; ++j) {
      const uint64_t q = q0 + (uint32_t)(j * kSortThreads + tid);
      bin[j] = K + 1;  // K = known query without a bin, K + 1 = past the end of the batch
      if (q < a.n_q) {
        const uint32_t g = qg[j];
        uint32_t gi = 0xffffu;
        if (direct) {
          if (g < (uint32_t)kBnDirectGroups) gi = s_gmap[g];
        } else {
          uint32_t lo = 0, hi = a.n_groups;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_grp[mid].gval < g) lo = mid + 1; else hi = mid;
          }
          if (lo < a.n_groups && s_grp[lo].gval == g) gi = lo;
        }
        bin[j] = K;
        if (gi != 0xffffu) {
          const uint32_t cell = ql[j] >> a.cell_shift;
          if (cell < s_grp[gi].n_cells) {  // bin = (number of bins that begin at or before the cell) - 1
            const uint32_t c = s_grp[gi].cell_base + cell;
            bin[j] = s_bits[n_words + (c >> 5)] + __popc(s_bits[c >> 5] & (0xffffffffu >> (31u - (c & 31u)))) - 1u;
          }
        }
      }
    }
    // ranks (issuing the eight match.any of a lane together, ahead of the counter updates, measured slower: 1.38 ms
    // against 1.18 ms for the whole kernel on config D)
#pragma unroll
    for (int j = 0; j < kSortQPT; ++j) {
      if (WARP_HIST) {
        const unsigned peers = __match_any_sync(0xffffffffu, bin[j]);  // lanes of this warp holding the same bin
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint16_t* cnt = s_whist + warp * Kp + bin[j];
        const uint32_t h = *cnt;
        __syncwarp();
        if (before == 0) *cnt = (uint16_t)(h + __popc(peers));
        __syncwarp();
        rank[j] = h + before;
      } else {
        rank[j] = bin[j] <= K ? atomicAdd(&s_hist[bin[j]], 1u) : 0u;
      }
    }
    __syncthreads();
    if (WARP_HIST)  // per bin: counts of the warps -> exclusive prefix over the warps, total into hist
      for (uint32_t b = tid; b < K + 2; b += kSortThreads) {
        uint32_t e = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
          const uint32_t t = s_whist[w * Kp + b];
          s_whist[w * Kp + b] = (uint16_t)e;
          e += t;
        }
        s_hist[b] = e;
      }
    if (WARP_HIST) __syncthreads();
    // exclusive scan of the K + 1 counters (contiguous pieces per thread), in place
    const uint32_t per = (K + 1 + kSortThreads - 1) / kSortThreads;
    const uint32_t b0 = min((uint32_t)tid * per, K + 1), b1 = min(b0 + per, K + 1);
    uint64_t mine = 0;
    for (uint32_t b = b0; b < b1; ++b) mine += s_hist[b];
    uint64_t total;
    uint64_t run = block_exclusive_scan<SumOp, kSortThreads>(mine, s_scan, &total);
    for (uint32_t b = b0; b < b1; ++b) {
      const uint32_t c = s_hist[b];
      s_hist[b] = (uint32_t)run;
      run += c;
    }
    __syncthreads();
    if (tid == 0) s_hist[K + 1] = (uint32_t)total;  // queries of the tile
    __syncthreads();
    uint16_t* trun = a.trun + (uint64_t)tile * (K + 2);
    for (uint32_t i = tid; i < K + 2; i += kSortThreads) trun[i] = (uint16_t)s_hist[i];
#pragma unroll
    for (int j = 0; j < kSortQPT; ++j)
      if (bin[j] <= K) {
        const uint32_t slot = s_hist[bin[j]] + rank[j] + (WARP_HIST ? (uint32_t)s_whist[warp * Kp + bin[j]] : 0u);
        BCU_DEV_ASSERT(slot < (uint32_t)kTileQ);
        s_rec[slot] = make_uint2(ql[j], qh[j]);
        s_loc[slot] = (uint16_t)(j * kSortThreads + tid);
      }
    __syncthreads();
    const uint32_t n_here = (uint32_t)total;
    for (uint32_t s = tid; s < n_here; s += kSortThreads) {
      a.brec[q0 + s] = s_rec[s];
      a.bloc[q0 + s] = s_loc[s];
    }
  }
}

// [tile][bin] starts -> [bin][tile] {start, length}; 32 x 32 patches through shared memory
__global__ void __launch_bounds__(256) bin_transpose_kernel(const BinnedArgs a) {
  __shared__ uint32_t s[32][33];
  const uint32_t K = a.n_bins;
  const uint32_t k0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const uint32_t t = t0 + r, k = k0 + tx;
    uint32_t v = 0;
    if (t < a.n_tiles && k < K) {
      const uint16_t* row = a.trun + (uint64_t)t * (K + 2);
      const uint32_t st = row[k], en = row[k + 1];
      v = st | ((en - st) << 16);
    }
    s[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const uint32_t k = k0 + r, t = t0 + tx;
    if (k < K && t < a.n_tiles) a.brun[(uint64_t)k * a.n_tiles + t] = s[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------------
// The tile probe. Dynamic shared memory: tile[kBinTileBytes] (own rows low | high | id, then the bin's blob:
// sub_start | cov_rel | cov_high | cov_id) | stage[kProbeWarps][kStageIds] u32.
// One CTA per SM (the tile fills its shared memory), so latency is hidden by the warps of that one CTA and by
// instruction-level parallelism inside a lane: candidate rows are tested four at a time, staged ids are fetched
// four at a time, and the queries of the step after next are brought into L2 (prefetch.global.L2) while the
// current step is being answered. Round 2 profile of the first version (length classes + max-length windows):
// issue-bound at 49 warp instructions per query, 58 % of them in the window test and the hit expansion -- hence
// the coverage lists (common.cuh): ~1.5 candidates per hit instead of 2.9 and no per-class work.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Shared memory through 32-bit shared-window addresses: with the kernel at its register cap the compiler
// otherwise rebuilds the generic base of the dynamic array before every access (4 extra instructions per load
// or store in the round-2 profile). The tile is read-only while a bin is being answered (plain asm, may be
// scheduled freely); the staging rows are written and read back by the same warp (volatile).
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// hit mask of the w <= 32 candidates whose `high` values start at shared address `addr`: bit j = (high[j] >= ql).
// Four candidates per 128-bit load from the 16-byte row the window starts in (the rows before the window are
// shifted out at the end; the arrays begin on a row and are followed by other arrays of the tile, so the first and
// the last load stay inside it).
__device__ __forceinline__ uint32_t window_mask(uint32_t addr, uint32_t w, uint32_t ql) {
  if (w == 0) return 0u;
  const uint32_t skip = (addr >> 2) & 3u;
  uint32_t row = addr & ~15u;
  uint64_t acc = 0;
  for (uint32_t k = 0; k < skip + w; k += 4, row += 16u) {
    const uint4 v = lds128(row);
    const uint32_t b = (uint32_t)(v.x >= ql) | ((uint32_t)(v.y >= ql) << 1) | ((uint32_t)(v.z >= ql) << 2) |
                       ((uint32_t)(v.w >= ql) << 3);
    acc |= (uint64_t)b << k;
  }
  return (uint32_t)(acc >> skip) & (w >= 32u ? 0xffffffffu : (1u << w) - 1u);
}

// the ids of the hits in `mask` (candidate j -> shared word ids + 4j) are appended at shared address `out`
__device__ __forceinline__ uint32_t expand_hits(uint32_t mask, uint32_t ids, uint32_t out) {
  while (mask) {
    const uint32_t j = (uint32_t)__ffs(mask) - 1u;
    mask &= mask - 1u;
    sts32(out, lds32(ids + 4u * j));
    out += 4u;
  }
  return out;
}
// the same for a warp whose hits exceed the staging row: only positions [0, kStageIds) of the round are kept
__device__ __forceinline__ uint32_t expand_hits_clipped(uint32_t mask, uint32_t ids, uint32_t stage, uint32_t p) {
  while (mask) {
    const uint32_t j = (uint32_t)__ffs(mask) - 1u;
    mask &= mask - 1u;
    if (p < (uint32_t)kStageIds) sts32(stage + 4u * p, lds32(ids + 4u * j));  // p wraps far above for earlier hits
    ++p;
  }
  return p;
}

template <bool EMIT>
__global__ void __launch_bounds__(kProbeThreads, 1) bin_probe_kernel(const BinnedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BinDesc s_desc;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_unit;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sm = smem_u32(smem_raw);
  const uint32_t a_stage = sm + kBinTileBytes + (uint32_t)warp * kStageIds * 4u;  // this warp's staging row
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t parity = 0;
  uint64_t slab_cur = 0, slab_end = 0;  // this warp's reserved piece of the staging area (warp-uniform)
  constexpr uint32_t kStride = kProbeWarps * 32;  // tiles between two steps of one warp

  for (;;) {
    __syncthreads();  // every warp is done with the previous bin's tile
    if (tid == 0) s_unit = atomicAdd(a.probe_ticket, 1u);
    __syncthreads();
    const uint32_t k = s_unit;
    if (k >= a.n_bins) break;
    for (uint32_t i = tid; i < sizeof(BinDesc) / 4; i += kProbeThreads)
      reinterpret_cast<uint32_t*>(&s_desc)[i] = reinterpret_cast<const uint32_t*>(a.desc + k)[i];
    __syncthreads();
    const uint32_t n_copy = s_desc.n_copy, nsub = s_desc.nsub, n_cov = s_desc.n_cov;
    const uint32_t tab = ((nsub + 1 + 7) & ~7u) * 2u;  // bytes of one u16 table
    // shared addresses of the tile's pieces: own rows low | high | id, then the blob sub_start | cov_rel | cov_high | cov_id
    const uint32_t a_low = sm, a_high = sm + 4u * n_copy, a_id = sm + 8u * n_copy, a_sub = sm + 12u * n_copy;
    const uint32_t a_rel = a_sub + tab, a_chigh = a_rel + tab, a_cid = a_chigh + 4u * n_cov;
    if (tid == 0) {  // one thread arms the barrier with the byte count and issues the bulk copies
      mbar_expect_tx(&s_bar, n_copy * 12u + s_desc.blob_bytes);
      if (n_copy) {
        bulk_g2s(smem_raw, a.low + s_desc.row0, n_copy * 4u, &s_bar);
        bulk_g2s(smem_raw + 4u * n_copy, a.high + s_desc.row0, n_copy * 4u, &s_bar);
        bulk_g2s(smem_raw + 8u * n_copy, a.ids + s_desc.row0, n_copy * 4u, &s_bar);
      }
      bulk_g2s(smem_raw + 12u * n_copy, a.blob + s_desc.blob, s_desc.blob_bytes, &s_bar);
    }
    // while the tile is on its way: this warp's first two run descriptors (and the L2 prefetch of their queries)
    const uint32_t* const runs = a.brun + (uint64_t)k * a.n_tiles;
    uint32_t t0 = warp * 32;
    uint32_t d0 = (t0 + lane < a.n_tiles) ? runs[t0 + lane] : 0u;
    uint32_t d1 = (t0 + kStride + lane < a.n_tiles) ? runs[t0 + kStride + lane] : 0u;
    if (d0 >> 16) prefetch_l2(a.brec + (uint64_t)(t0 + lane) * kTileQ + (d0 & 0xffffu));
    if (d1 >> 16) prefetch_l2(a.brec + (uint64_t)(t0 + kStride + lane) * kTileQ + (d1 & 0xffffu));
    mbar_wait(&s_bar, parity);
    parity ^= 1u;

    const uint32_t x_begin = s_desc.x_begin, x_end = s_desc.x_end, ls = s_desc.ls;
    for (; t0 < a.n_tiles; t0 += kStride) {  // 32 tiles' runs of this bin per warp step
      // two steps ahead: the run descriptors (a plain load, consumed at the end of this step)
      const uint32_t t2 = t0 + 2 * kStride + lane;
      const uint32_t d2 = t2 < a.n_tiles ? runs[t2] : 0u;
      const uint32_t r_start = d0 & 0xffffu, r_n = d0 >> 16;
      const uint32_t r_incl = warp_incl_scan(r_n, lane);
      const uint32_t m = __shfl_sync(0xffffffffu, r_incl, 31);
      for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t f = base + lane;
        const bool valid = f < m;
        const int src = run_owner(r_incl, valid ? f : 0u);
        const uint32_t o_incl = __shfl_sync(0xffffffffu, r_incl, src), o_n = __shfl_sync(0xffffffffu, r_n, src);
        const uint32_t o_start = __shfl_sync(0xffffffffu, r_start, src);
        const uint64_t slot = (uint64_t)(t0 + src) * kTileQ + o_start + (f - (o_incl - o_n));
        uint2 q = make_uint2(x_begin, x_begin);
        BCU_DEV_ASSERT(!valid || slot < (uint64_t)a.n_tiles * kTileQ);
        if (valid) q = a.brec[slot];
        const uint32_t ql = q.x, qh = q.y;

        // candidates: the coverage list of q.low's sub-cell + the own rows from that sub-cell's start up to q.high
        const uint32_t g = (ql - x_begin) >> ls;  // < nsub: the query was routed here by q.low
        BCU_DEV_ASSERT(ql >= x_begin && g < nsub);
        const uint32_t c0 = lds16(a_rel + 2u * g), wa = lds16(a_rel + 2u * g + 2u) - c0;
        const uint32_t sa = lds16(a_sub + 2u * g);
        uint32_t ub = lds16(a_sub + 2u * min((max(qh, ql) - x_begin) >> ls, nsub - 1u) + 2u);
        while (ub > sa && lds32(a_low + 4u * ub - 4u) > qh) --ub;  // back over the own rows that start beyond q.high
        const uint32_t wb = ub > sa ? ub - sa : 0u;
        BCU_DEV_ASSERT(c0 + wa <= n_cov && ub <= n_copy && sa <= n_copy && 12u * n_copy + s_desc.blob_bytes <= kBinTileBytes);
        // the general index answers: inverted queries, queries that reach past the bin, windows beyond the mask
        const bool spill = valid && (qh < ql || (x_end != 0 && qh >= x_end) || wa > kMaskRows || wb > kMaskRows);
        uint32_t mask_a = 0, mask_b = 0;
        if (valid && !spill) {
          mask_a = window_mask(a_chigh + 4u * c0, wa, ql);
          mask_b = window_mask(a_high + 4u * sa, wb, ql);
        }
        const uint32_t cnt = __popc(mask_a) + __popc(mask_b);
        if (spill) {
          const uint32_t at = atomicAdd(a.n_spill, 1u);
          a.spill[2ull * at] = (uint32_t)slot;  // slots stay below 2^32 (checked on the host)
          a.spill[2ull * at + 1] = k;
        }
        if (!EMIT) {
          if (valid && !spill) a.sres[slot] = make_uint2(0u, cnt);
          continue;
        }
        // ---- compact the warp's hits: ranks -> staging reservation -> shared-memory expansion -> coalesced store
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t excl = incl - cnt;
        uint64_t wbase = 0;
        if (total) {
          if (slab_cur + total > slab_end) {
            uint64_t got = 0;
            const uint32_t want = max(total, kSlabIds);
            if (lane == 0) got = atomicAdd(a.stage_cursor, (unsigned long long)want);
            got = shfl_u64(got, 0);
            slab_cur = got;
            slab_end = got + want;
          }
          wbase = slab_cur;
          slab_cur += total;
        }
        const bool fits = wbase + total <= a.stage_cap;  // otherwise the join exceeds the caller's pair capacity
        if (valid && !spill) a.sres[slot] = make_uint2(fits ? (uint32_t)(wbase + excl) : kNotStored, cnt);
        if (!fits) continue;
        uint32_t* const out = a.staging + wbase;
        BCU_DEV_ASSERT(wbase + total <= a.stage_cap && excl + cnt <= total);
        if (total <= (uint32_t)kStageIds) {  // the usual case: one round, no clipping
          const uint32_t p = expand_hits(mask_a, a_cid + 4u * c0, a_stage + 4u * excl);
          expand_hits(mask_b, a_id + 4u * sa, p);
          __syncwarp();
#pragma unroll
          for (uint32_t s = 0; s < (uint32_t)kStageIds; s += 32)
            if (s + lane < total) out[s + lane] = lds32_volatile(a_stage + 4u * (s + lane));
          __syncwarp();
        } else {
          for (uint32_t r0 = 0; r0 < total; r0 += kStageIds) {
            const uint32_t p = expand_hits_clipped(mask_a, a_cid + 4u * c0, a_stage, excl - r0);
            expand_hits_clipped(mask_b, a_id + 4u * sa, a_stage, p);
            __syncwarp();
            const uint32_t n_here = min((uint32_t)kStageIds, total - r0);
            for (uint32_t s = lane; s < n_here; s += 32) out[r0 + s] = lds32_volatile(a_stage + 4u * s);
            __syncwarp();
          }
        }
      }
      // rotate the descriptors; bring the queries of the step after next into L2 (first and last sector of each run)
      d0 = d1;
      d1 = d2;
      if (d2 >> 16) {
        const uint2* qrun = a.brec + (uint64_t)t2 * kTileQ + (d2 & 0xffffu);
        prefetch_l2(qrun);
        prefetch_l2(qrun + (d2 >> 16) - 1);
      }
    }
  }
}

// The queries the tile probe gave up on (window beyond the hit mask, or reaching past their bin): one warp per
// query through the general index -- directory bounds (join.cu query_bounds) + an exact scan of the rows.
template <bool EMIT>
__global__ void __launch_bounds__(256) bin_spill_kernel(const BinnedArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = *a.n_spill;
  for (uint32_t e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += gridDim.x * 8) {
    const uint32_t slot = a.spill[2ull * e], g = a.desc[a.spill[2ull * e + 1]].group;
    const uint2 q = a.brec[slot];
    uint32_t lbs[4], ubs[4];
    uint32_t cnt = 0;
    for (uint32_t c = 0; c < a.n_cls; ++c) {
      const GroupDesc d = a.gtable[(size_t)c * a.n_groups + g];
      lbs[c] = ubs[c] = 0;
      const uint32_t b_lo = q.x >> d.shift;
      if (b_lo >= d.nb) continue;
      const uint32_t b_hi = min(q.y >> d.shift, d.nb - 1u);
      const uint32_t lb = a.dir[d.bin_base + b_lo].lb, ub = a.dir[d.bin_base + b_hi].ub;
      if (ub <= lb) continue;
      lbs[c] = lb;
      ubs[c] = ub;
      for (uint32_t r = lb + lane; r < ub; r += 32) {
        const uint2 t = a.lowhigh[r];
        cnt += (uint32_t)((q.x <= t.y) & (t.x <= q.y));
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    uint64_t base = 0;
    if (EMIT && cnt) {
      if (lane == 0) base = atomicAdd(a.stage_cursor, (unsigned long long)cnt);
      base = shfl_u64(base, 0);
    }
    const bool fits = base + cnt <= a.stage_cap;
    if (lane == 0) a.sres[slot] = make_uint2(fits ? (uint32_t)base : kNotStored, cnt);
    if (!EMIT || cnt == 0 || !fits) continue;
    uint32_t done = 0;
    for (uint32_t c = 0; c < a.n_cls; ++c)
      for (uint32_t r0 = lbs[c]; r0 < ubs[c]; r0 += 32) {  // warp-uniform bounds
        const uint32_t r = r0 + lane;
        bool hit = false;
        if (r < ubs[c]) {
          const uint2 t = a.lowhigh[r];
          hit = (q.x <= t.y) & (t.x <= q.y);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) a.staging[base + done + __popc(bal & ((1u << lane) - 1u))] = a.gids[r];
        done += __popc(bal);
      }
  }
}

// ---------------------------------------------------------------------------------------------------
// Back to query order. Dynamic shared memory: off[kTileQ + 2] u64 | out[kPlaceCap] u32 | loc[kTileQ] u16 (two
// CTAs per SM). Offsets inside a tile are 64-bit: a spilled query may hit every target.
template <bool EMIT>
__global__ void __launch_bounds__(kPlaceThreads, 2) bin_place_kernel(const BinnedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_off = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* s_out = reinterpret_cast<uint32_t*>(s_off + kTileQ + 2);
  uint16_t* s_loc = reinterpret_cast<uint16_t*>(s_out + kPlaceCap + 4);
  __shared__ uint64_t s_scan[kPlaceThreads / 32 + 1];
  __shared__ uint32_t s_tile;
  __shared__ uint64_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(a.place_ticket, 1u);  // ticket order: every earlier tile has started (look-back)
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t K = a.n_bins;
  const uint16_t* trun = a.trun + (uint64_t)tile * (K + 2);
  const uint64_t q0 = (uint64_t)tile * kTileQ;
  const uint2* const sres = a.sres + q0;
  // Everything the tile needs from global memory is requested at once (one round trip instead of three dependent
  // ones), and the tile's hit total is published for the look-back of the later tiles as soon as the counts are
  // there -- before the scatter and the scan: in the round-2 profile a third of this phase was the look-back warp
  // spinning on predecessors that had not published yet (and the other warps waiting for it at the barrier).
  const uint32_t n_here = (uint32_t)min((uint64_t)kTileQ, (uint64_t)a.n_q - q0);  // queries of the tile (== trun[K + 1])
  const uint32_t n_live_ld = trun[K];                                             // slots with a bin
  uint32_t cnt_r[kPlaceQPT], loc_r[kPlaceQPT];
#pragma unroll
  for (int k = 0; k < kPlaceQPT; ++k) {
    const uint32_t s = (uint32_t)(k * kPlaceThreads + tid);
    loc_r[k] = s < n_here ? (uint32_t)a.bloc[q0 + s] : 0u;
    cnt_r[k] = s < n_here ? sres[s].y : 0u;  // (slots without a bin hold no result: masked below)
    s_off[s] = 0;
  }
  const uint32_t n_live = n_live_ld;
  BCU_DEV_ASSERT(n_live <= n_here && n_here <= (uint32_t)kTileQ && n_here == trun[K + 1]);
  uint64_t early = 0;
#pragma unroll
  for (int k = 0; k < kPlaceQPT; ++k) {
    if ((uint32_t)(k * kPlaceThreads + tid) >= n_live) cnt_r[k] = 0;
    early += cnt_r[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) early += shfl_down_u64(early, o);
  if (lane == 0) s_scan[warp] = early;
  __syncthreads();  // (also: s_off is zeroed)
  if (warp == 0) {
    uint64_t v = lane < kPlaceThreads / 32 ? s_scan[lane] : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += shfl_down_u64(v, o);
    if (lane == 0) reinterpret_cast<volatile uint64_t*>(a.status)[tile] = (tile == 0 ? kStInclusive : kStAggregate) | v;
  }
#pragma unroll
  for (int k = 0; k < kPlaceQPT; ++k) {  // counts by query position
    const uint32_t s = (uint32_t)(k * kPlaceThreads + tid);
    s_loc[s] = (uint16_t)loc_r[k];
    if (s < n_live) s_off[loc_r[k]] = cnt_r[k];
  }
  __syncthreads();
  // exclusive scan over the tile's 4096 counts (kPlaceQPT consecutive entries per thread)
  uint32_t c[kPlaceQPT];
  uint64_t mine = 0;
#pragma unroll
  for (int j = 0; j < kPlaceQPT; ++j) { c[j] = (uint32_t)s_off[tid * kPlaceQPT + j]; mine += c[j]; }
  uint64_t tile_total;
  uint64_t run = block_exclusive_scan<SumOp, kPlaceThreads>(mine, s_scan, &tile_total);
#pragma unroll
  for (int j = 0; j < kPlaceQPT; ++j) { s_off[tid * kPlaceQPT + j] = run; run += c[j]; }
  if (tid == 0) s_off[kTileQ] = tile_total;
  if (warp == 0) {
    const uint64_t e = lookback_exclusive<SumOp>(a.status, tile, tile_total);
    if (lane == 0) s_base = e;
  }
  __syncthreads();
  const uint64_t base = s_base;
  for (uint32_t i = tid; i < n_here; i += kPlaceThreads) a.offsets[q0 + i] = base + s_off[i];
  if (tile == a.n_tiles - 1 && tid == 0) {
    const uint64_t t = base + tile_total;
    a.offsets[a.n_q] = t;
    if (a.total) *a.total = t;
    if (a.total_mapped) *a.total_mapped = t;
  }
  if (!EMIT) return;

  const uint64_t total = tile_total;
  // window position p lives in s_out[p + shift], shift = (first output position of the window) mod 4, so that
  // 16-byte rows of shared memory line up with 16-byte rows of the output columns
  auto write_out = [&](uint32_t* __restrict__ col, bool aligned, uint64_t first, uint32_t shift, uint32_t n_out) {
    const uint64_t row0 = first - shift;  // multiple of 4
    const uint32_t end = n_out + shift;
    for (uint32_t p4 = tid * 4; p4 < end; p4 += kPlaceThreads * 4) {
      const uint64_t pos = row0 + p4;
      if (aligned && p4 >= shift && p4 + 4 <= end && pos + 4 <= a.capacity) {
        *reinterpret_cast<uint4*>(col + pos) = *reinterpret_cast<const uint4*>(s_out + p4);
      } else {
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u)
          if (p4 + u >= shift && p4 + u < end && pos + u < a.capacity) col[pos + u] = s_out[p4 + u];
      }
    }
  };
  for (uint64_t w0 = 0; w0 < total; w0 += kPlaceCap) {
    const uint32_t shift = (uint32_t)((base + w0) & 3u);
    // gather: a lane per slot copies its own id list (the lists of neighbouring slots lie next to each other in
    // the staging area, so the lanes of a warp read the same few sectors; four loads are issued before the stores).
    // Measured against warp-cooperative copies of the flattened lists (owner of every id by shuffle search, or by
    // a bitmap of list ends + find-nth-set): those cost 1.8x - 2x the instructions of this loop.
    for (uint32_t s0 = warp * 32; s0 < n_live; s0 += kPlaceThreads) {
      const uint32_t s = s0 + lane;
      uint2 r = make_uint2(0u, 0u);
      uint64_t dst0 = 0;
      if (s < n_live) { r = sres[s]; dst0 = s_off[s_loc[s]]; }
      // the part [j_lo, j_hi) of the list that falls into the window, once, in 64 bits; the copy loop is 32-bit
      const uint64_t w1 = w0 + kPlaceCap;
      const uint32_t j_lo = dst0 >= w0 ? 0u : (uint32_t)min((uint64_t)r.y, w0 - dst0);
      const uint32_t j_hi = w1 > dst0 ? (uint32_t)min((uint64_t)r.y, w1 - dst0) : 0u;
      const uint32_t n = j_hi > j_lo ? j_hi - j_lo : 0u;
      const uint32_t steps = __reduce_max_sync(0xffffffffu, n);
      if (steps == 0) continue;
      const bool stored = r.x != kNotStored;
      BCU_DEV_ASSERT(!(n && stored) || (uint64_t)r.x + j_hi <= a.stage_cap);
      const uint32_t* sp = a.staging + r.x + j_lo;
      uint32_t* op = s_out + shift + (uint32_t)(dst0 + j_lo - w0);
      for (uint32_t j0 = 0; j0 < steps; j0 += kPlaceLoads) {
        uint32_t v[kPlaceLoads];
#pragma unroll
        for (uint32_t u = 0; u < kPlaceLoads; ++u) v[u] = (j0 + u < n && stored) ? sp[j0 + u] : kNotStored;
#pragma unroll
        for (uint32_t u = 0; u < kPlaceLoads; ++u)
          if (j0 + u < n) op[j0 + u] = v[u];
      }
    }
    __syncthreads();
    const uint32_t n_out = (uint32_t)min((uint64_t)kPlaceCap, total - w0);
    write_out(a.hit_target, a.ht_aligned != 0, base + w0, shift, n_out);
    __syncthreads();
    if (a.hit_query) {  // the query-id column: every query fills its own stretch of the window
      for (uint32_t i = tid; i < n_here; i += kPlaceThreads) {
        const uint64_t b = max(s_off[i], w0), e = min(s_off[i + 1], w0 + (uint64_t)kPlaceCap);
        const uint32_t qid = a.qid_base + (uint32_t)(q0 + i);
        if (e > b) {
          uint32_t* op = s_out + shift + (uint32_t)(b - w0);
          const uint32_t n = (uint32_t)(e - b);
          for (uint32_t p = 0; p < n; ++p) op[p] = qid;
        }
      }
      __syncthreads();
      write_out(a.hit_query, a.hq_aligned != 0, base + w0, shift, n_out);
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

int launch_join_binned(const bcu_index* ix, int mode, uint64_t n_q, const uint32_t* d_qgroup, const uint32_t* d_qlow,
                       const uint32_t* d_qhigh, uint64_t* d_offsets, uint64_t pair_capacity, uint32_t* d_hit_query,
                       uint32_t* d_hit_target, uint64_t* d_total, uint32_t query_id_base, cudaStream_t stream,
                       uint64_t* total_mapped) {
  // eligibility: an index with a bin layout, a batch large enough to pay for the routing pass (BCU_BINNED=1
  // forces the path, =0 disables it), staging addressable with 32 bits
  const int force = env_int("BCU_BINNED", -1);
  if (ix->bn_bins == 0 || force == 0 || (mode != kModeFused && mode != kModeCount)) return BCU_NOT_TAKEN;
  if (ix->n_groups > (uint32_t)kMaxSmemGroups) return BCU_NOT_TAKEN;
  const uint64_t min_q = (uint64_t)env_int("BCU_BINNED_MIN_QUERIES", 1 << 21);
  const uint64_t min_bytes = (uint64_t)env_int("BCU_BINNED_MIN_INDEX_MB", 96) << 20;
  if (force != 1 && (n_q < min_q || ix->bytes < min_bytes)) return BCU_NOT_TAKEN;
  if (n_q == 0 || n_q > 0xffffffffull - 2 * kTileQ) return BCU_NOT_TAKEN;
  const bool emit = mode == kModeFused;
  const uint64_t n_tiles = (n_q + kTileQ - 1) / kTileQ;
  const int sms = [&] {
    int v = 0;
    return (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, ix->device) == cudaSuccess && v > 0) ? v : 148;
  }();
  const uint64_t slack = (uint64_t)sms * kProbeWarps * kSlabIds + (1u << 20);
  const uint64_t stage_cap = emit ? pair_capacity + slack : 0;
  if (stage_cap >= 0xffffffffull) return BCU_NOT_TAKEN;

  BinnedArgs a;
  a.low = ix->d_bn_low; a.high = ix->d_bn_high; a.ids = ix->d_bn_id; a.blob = ix->d_bn_blob;
  a.lowhigh = ix->d_lowhigh; a.gids = ix->d_id; a.dir = ix->d_dir;
  a.gtable = ix->d_groups; a.desc = ix->d_bn_desc; a.groups = ix->d_bn_groups; a.cellbits = ix->d_bn_cellbits;
  a.n_groups = ix->n_groups; a.n_bins = ix->bn_bins; a.n_cells = ix->bn_cells; a.cell_shift = ix->bn_cell_shift;
  a.max_gval = ix->max_gval; a.n_cls = ix->n_comp;
  a.qgroup = d_qgroup; a.qlow = d_qlow; a.qhigh = d_qhigh; a.n_q = (uint32_t)n_q; a.n_tiles = (uint32_t)n_tiles;
  a.offsets = d_offsets; a.capacity = pair_capacity; a.hit_query = d_hit_query; a.hit_target = d_hit_target;
  a.total = d_total; a.total_mapped = total_mapped; a.qid_base = query_id_base; a.emit = emit;
  a.stage_cap = stage_cap;
  a.ht_aligned = (reinterpret_cast<uintptr_t>(d_hit_target) & 15u) == 0;
  a.hq_aligned = (reinterpret_cast<uintptr_t>(d_hit_query) & 15u) == 0;

  // one stream-ordered allocation, carved up (every piece 256-byte aligned)
  const uint32_t K = ix->bn_bins;
  const uint64_t slots = n_tiles * kTileQ;
  uint64_t at = 0;
  auto carve = [&](uint64_t bytes) { const uint64_t o = at; at += (bytes + 255) / 256 * 256; return o; };
  const uint64_t o_ctr = carve(256), o_status = carve(n_tiles * 8), o_brec = carve(slots * 8), o_sres = carve(slots * 8),
                 o_bloc = carve(slots * 2), o_trun = carve(n_tiles * (K + 2) * 2), o_brun = carve((uint64_t)K * n_tiles * 4),
                 o_spill = carve(n_q * 8), o_stage = carve(stage_cap * 4);
  char* scratch = nullptr;
  BCU_CUDA(cudaMallocAsync((void**)&scratch, at, stream));
  struct ScratchGuard {
    void* p;
    cudaStream_t s;
    ~ScratchGuard() { cudaFreeAsync(p, s); }
  } guard{scratch, stream};
  BCU_CUDA(cudaMemsetAsync(scratch, 0, o_brec, stream));  // counters + look-back status words
  a.stage_cursor = reinterpret_cast<unsigned long long*>(scratch + o_ctr);
  a.probe_ticket = reinterpret_cast<uint32_t*>(scratch + o_ctr + 8);
  a.n_spill = reinterpret_cast<uint32_t*>(scratch + o_ctr + 12);
  a.place_ticket = reinterpret_cast<uint32_t*>(scratch + o_ctr + 16);
  a.status = reinterpret_cast<uint64_t*>(scratch + o_status);
  a.brec = reinterpret_cast<uint2*>(scratch + o_brec);
  a.sres = reinterpret_cast<uint2*>(scratch + o_sres);
  a.bloc = reinterpret_cast<uint16_t*>(scratch + o_bloc);
  a.trun = reinterpret_cast<uint16_t*>(scratch + o_trun);
  a.brun = reinterpret_cast<uint32_t*>(scratch + o_brun);
  a.spill = reinterpret_cast<uint32_t*>(scratch + o_spill);
  a.staging = reinterpret_cast<uint32_t*>(scratch + o_stage);

  const bool warp_hist = K <= kSortMaxWarpHistBins;
  const size_t sort_smem = (size_t)kTileQ * 8 + (size_t)((K + 2 + 3) & ~3u) * 4 + (size_t)ix->n_groups * sizeof(BinGroup) +
                           (size_t)kTileQ * 2 + (size_t)((2 * ((ix->bn_cells + 31) / 32) + 3) & ~3u) * 4 + (size_t)kBnDirectGroups * 2 +
                           (warp_hist ? (size_t)kSortWarps * ((K + 2 + 3) & ~3u) * 2 : 0);
  const size_t probe_smem = (size_t)kBinTileBytes + (size_t)kProbeWarps * kStageIds * 4;
  const size_t place_smem = (size_t)(kTileQ + 2) * 8 + (size_t)(kPlaceCap + 4) * 4 + (size_t)kTileQ * 2;
  static std::atomic<bool> attrs_set[64];
  if (ix->device < 0 || ix->device >= 64 || !attrs_set[ix->device].load()) {
    BCU_CUDA(cudaFuncSetAttribute(bin_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BCU_CUDA(cudaFuncSetAttribute(bin_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BCU_CUDA(cudaFuncSetAttribute(bin_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem));
    BCU_CUDA(cudaFuncSetAttribute(bin_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem));
    BCU_CUDA(cudaFuncSetAttribute(bin_place_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
    BCU_CUDA(cudaFuncSetAttribute(bin_place_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
    if (ix->device >= 0 && ix->device < 64) attrs_set[ix->device].store(true);
  }
  if (sort_smem > 200 * 1024) return BCU_NOT_TAKEN;

  const unsigned sort_grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sms * std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (sort_smem + 1024))));
  if (warp_hist) bin_sort_kernel<true><<<sort_grid, kSortThreads, sort_smem, stream>>>(a);
  else bin_sort_kernel<false><<<sort_grid, kSortThreads, sort_smem, stream>>>(a);
  BCU_LAUNCHED();
  bin_transpose_kernel<<<dim3((K + 31) / 32, (unsigned)((n_tiles + 31) / 32)), 256, 0, stream>>>(a);
  BCU_LAUNCHED();
  const unsigned probe_grid = (unsigned)std::min<uint32_t>(K, (uint32_t)sms);
  if (emit) bin_probe_kernel<true><<<probe_grid, kProbeThreads, probe_smem, stream>>>(a);
  else bin_probe_kernel<false><<<probe_grid, kProbeThreads, probe_smem, stream>>>(a);
  BCU_LAUNCHED();
  if (emit) bin_spill_kernel<true><<<sms * 4, 256, 0, stream>>>(a);
  else bin_spill_kernel<false><<<sms * 4, 256, 0, stream>>>(a);
  BCU_LAUNCHED();
  if (emit) bin_place_kernel<true><<<(unsigned)n_tiles, kPlaceThreads, place_smem, stream>>>(a);
  else bin_place_kernel<false><<<(unsigned)n_tiles, kPlaceThreads, place_smem, stream>>>(a);
  BCU_LAUNCHED();
  return BCU_OK;
}

}  // namespace bcu
