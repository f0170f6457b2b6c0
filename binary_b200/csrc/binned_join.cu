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
constexpr int kProbeThreads = 512;
constexpr int kProbeWarps = kProbeThreads / 32;
constexpr int kStageIds = 512;        // hits staged per warp and round
constexpr uint32_t kSlabIds = 8192;   // staging is reserved per warp in slabs: one global atomic per ~40 rounds
constexpr int kPlaceThreads = 512;
constexpr int kPlaceQPT = kTileQ / kPlaceThreads;
constexpr int kPlaceCap = 32768;      // target ids assembled per tile and round
constexpr uint32_t kMaskRows = 32;    // candidate window a lane can record (one hit-mask word per class)
constexpr int kBnDirectGroups = 1024; // group values below this are routed through a direct map
constexpr uint32_t kNotStored = 0xffffffffu;  // sres.x of a hit list that did not fit the staging area (the join
                                              // exceeds the caller's pair capacity; staging indices stay below it)

struct BinnedArgs {
  // index
  const uint32_t* __restrict__ low;
  const uint32_t* __restrict__ high;
  const uint32_t* __restrict__ ids;
  const uint2* __restrict__ lowhigh;
  const DirEntry* __restrict__ dir;
  const GroupDesc* __restrict__ gtable;  // [n_cls][n_groups] (general index: used by the spill path)
  const BinDesc* __restrict__ desc;
  const BinGroup* __restrict__ groups;
  const uint16_t* __restrict__ cell2bin;
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
// Dynamic shared memory: hist[n_bins + 2] u32 | rec[kTileQ] uint2 | loc[kTileQ] u16 | cell2bin[n_cells] u16 |
// groups[n_groups] BinGroup | gmap[kBnDirectGroups] u16
__global__ void __launch_bounds__(kSortThreads) bin_sort_kernel(const BinnedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t K = a.n_bins;
  uint2* s_rec = reinterpret_cast<uint2*>(smem_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_rec + kTileQ);
  BinGroup* s_grp = reinterpret_cast<BinGroup*>(s_hist + ((K + 2 + 3) & ~3u));
  uint16_t* s_loc = reinterpret_cast<uint16_t*>(s_grp + a.n_groups);
  uint16_t* s_c2b = s_loc + kTileQ;
  uint16_t* s_gmap = s_c2b + ((a.n_cells + 7) & ~7u);
  __shared__ uint64_t s_scan[kSortThreads / 32 + 1];
  const int tid = threadIdx.x;
  const bool direct = a.max_gval < (uint32_t)kBnDirectGroups;

  for (uint32_t i = tid; i < a.n_cells; i += kSortThreads) s_c2b[i] = a.cell2bin[i];
  for (uint32_t i = tid; i < a.n_groups; i += kSortThreads) s_grp[i] = a.groups[i];
  if (direct)
    for (int i = tid; i < kBnDirectGroups; i += kSortThreads) s_gmap[i] = 0xffffu;
  __syncthreads();
  if (direct)
    for (uint32_t i = tid; i < a.n_groups; i += kSortThreads) s_gmap[s_grp[i].gval] = (uint16_t)i;

  for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    __syncthreads();  // previous tile's copy-out is done; the group map is visible
    for (uint32_t i = tid; i < K + 2; i += kSortThreads) s_hist[i] = 0;
    __syncthreads();
    const uint64_t q0 = (uint64_t)tile * kTileQ;
    uint32_t ql[kSortQPT], qh[kSortQPT], bin[kSortQPT], rank[kSortQPT];
#pragma unroll
    for (int j = 0; j < kSortQPT; ++j) {
      const uint64_t q = q0 + (uint32_t)(j * kSortThreads + tid);
      bin[j] = K + 1;  // K = known query without a bin, K + 1 = past the end of the batch
      ql[j] = qh[j] = 0;
      if (q < a.n_q) {
        ql[j] = a.qlow[q];
        qh[j] = a.qhigh[q];
        const uint32_t g = a.qgroup ? a.qgroup[q] : 0u;
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
          if (cell < s_grp[gi].n_cells) {
            const uint32_t b = s_c2b[s_grp[gi].cell_base + cell];
            if (b != kBinNull) bin[j] = b;
          }
        }
      }
      rank[j] = bin[j] <= K ? atomicAdd(&s_hist[bin[j]], 1u) : 0u;
    }
    __syncthreads();
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
    if (tid == 0) s_hist[K + 1] = (uint32_t)total;  // queries of the tile
    __syncthreads();
    uint16_t* trun = a.trun + (uint64_t)tile * (K + 2);
    for (uint32_t i = tid; i < K + 2; i += kSortThreads) trun[i] = (uint16_t)s_hist[i];
#pragma unroll
    for (int j = 0; j < kSortQPT; ++j)
      if (bin[j] <= K) {
        const uint32_t slot = s_hist[bin[j]] + rank[j];
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
// The tile probe. Dynamic shared memory: low[kBinRowsCap] | high[kBinRowsCap] | id[kBinRowsCap] (u32) |
// stage[kProbeWarps][kStageIds] u32 | lut[kBinLutCap] u16
struct ProbeHit {
  uint32_t mask[kBinMaxClasses];  // bit j = row lb + j of the class's window is a hit
  uint32_t lb[kBinMaxClasses];    // tile-relative first row of the window (already offset by the class's s_off)
};

template <bool EMIT>
__global__ void __launch_bounds__(kProbeThreads, 1) bin_probe_kernel(const BinnedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_low = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_high = s_low + kBinRowsCap;
  uint32_t* s_id = s_high + kBinRowsCap;
  uint32_t* s_stage = s_id + kBinRowsCap;
  uint16_t* s_lut = reinterpret_cast<uint16_t*>(s_stage + kProbeWarps * kStageIds);
  __shared__ BinDesc s_desc;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_unit;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* const my_stage = s_stage + warp * kStageIds;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t parity = 0;
  uint64_t slab_cur = 0, slab_end = 0;  // this warp's reserved piece of the staging area (warp-uniform)

  for (;;) {
    __syncthreads();  // every warp is done with the previous bin's tile
    if (tid == 0) s_unit = atomicAdd(a.probe_ticket, 1u);
    __syncthreads();
    const uint32_t k = s_unit;
    if (k >= a.n_bins) break;
    for (uint32_t i = tid; i < sizeof(BinDesc) / 4; i += kProbeThreads)
      reinterpret_cast<uint32_t*>(&s_desc)[i] = reinterpret_cast<const uint32_t*>(a.desc + k)[i];
    __syncthreads();
    if (tid == 0) {  // one thread arms the barrier with the byte count and issues the bulk copies
      mbar_expect_tx(&s_bar, s_desc.n_rows * 12u);
      for (uint32_t c = 0; c < a.n_cls; ++c) {
        const BinClass& kc = s_desc.cls[c];
        if (kc.n_copy == 0) continue;
        bulk_g2s(s_low + kc.s_off, a.low + kc.row0, kc.n_copy * 4u, &s_bar);
        bulk_g2s(s_high + kc.s_off, a.high + kc.row0, kc.n_copy * 4u, &s_bar);
        bulk_g2s(s_id + kc.s_off, a.ids + kc.row0, kc.n_copy * 4u, &s_bar);
      }
    }
    mbar_wait(&s_bar, parity);
    parity ^= 1u;
    // sub-cell tables: lut[j] = first row of the class (tile-relative, in [lo, hi]) with low >= x0 + (j << ls)
    for (uint32_t c = 0; c < a.n_cls; ++c) {
      const BinClass kc = s_desc.cls[c];
      uint16_t* lut = s_lut + kc.lut_off;
      if (kc.hi <= kc.lo) {
        for (uint32_t j = tid; j <= kc.nsub; j += kProbeThreads) lut[j] = (uint16_t)kc.lo;
        continue;
      }
      for (uint32_t r = kc.lo + tid; r < kc.hi; r += kProbeThreads) {
        const uint32_t cell = (s_low[kc.s_off + r] - kc.x0) >> kc.ls;               // < nsub
        const int64_t prev = r > kc.lo ? (int64_t)((s_low[kc.s_off + r - 1] - kc.x0) >> kc.ls) : -1;
        for (int64_t j = prev + 1; j <= (int64_t)cell; ++j) lut[j] = (uint16_t)r;
        if (r == kc.hi - 1)
          for (uint32_t j = cell + 1; j <= kc.nsub; ++j) lut[j] = (uint16_t)kc.hi;
      }
    }
    __syncthreads();

    const uint32_t x_end = s_desc.x_end;
    const uint32_t* const runs = a.brun + (uint64_t)k * a.n_tiles;
    for (uint32_t t0 = warp * 32; t0 < a.n_tiles; t0 += kProbeWarps * 32) {  // 32 tiles' runs of this bin per warp step
      const uint32_t t = t0 + lane;
      const uint32_t run = t < a.n_tiles ? runs[t] : 0u;
      const uint32_t r_start = run & 0xffffu, r_n = run >> 16;
      const uint32_t r_incl = warp_incl_scan(r_n, lane);
      const uint32_t m = __shfl_sync(0xffffffffu, r_incl, 31);
      for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t f = base + lane;
        const bool valid = f < m;
        const int src = run_owner(r_incl, valid ? f : 0u);
        const uint32_t o_incl = __shfl_sync(0xffffffffu, r_incl, src), o_n = __shfl_sync(0xffffffffu, r_n, src);
        const uint32_t o_start = __shfl_sync(0xffffffffu, r_start, src);
        const uint64_t slot = (uint64_t)(t0 + src) * kTileQ + o_start + (f - (o_incl - o_n));
        uint2 q = make_uint2(0, 0);
        if (valid) q = a.brec[slot];
        const uint32_t ql = q.x, qh = q.y;

        ProbeHit h;
        uint32_t cnt = 0;
        bool spill = valid && x_end != 0 && qh >= x_end && qh >= ql;  // reaches past the bin: rows beyond the tile
#pragma unroll
        for (uint32_t c = 0; c < kBinMaxClasses; ++c) {
          h.mask[c] = 0;
          h.lb[c] = 0;
          if (c >= a.n_cls || !valid || spill) continue;
          const BinClass& kc = s_desc.cls[c];
          if (kc.hi <= kc.lo || qh < kc.x0) continue;  // every row of the tile starts at or after x0
          const uint32_t* low = s_low + kc.s_off;
          const uint16_t* lut = s_lut + kc.lut_off;
          uint32_t ub = lut[min((qh - kc.x0) >> kc.ls, kc.nsub - 1u) + 1u];
          while (ub > kc.lo && low[ub - 1] > qh) --ub;            // first row with low > q.high
          const uint32_t lob = ql > kc.maxlen ? ql - kc.maxlen : 0u;  // rows starting before it cannot reach q.low
          uint32_t lb = kc.lo;
          if (lob > kc.x0) {
            lb = lut[min((lob - kc.x0) >> kc.ls, kc.nsub - 1u)];
            while (lb < ub && low[lb] < lob) ++lb;
          }
          const uint32_t w = ub > lb ? ub - lb : 0u;
          if (w > kMaskRows) { spill = true; continue; }
          const uint32_t* high = s_high + kc.s_off + lb;
          uint32_t mask = 0;
          for (uint32_t j = 0; j < w; ++j) mask |= (uint32_t)(high[j] >= ql) << j;
          h.mask[c] = mask;
          h.lb[c] = kc.s_off + lb;
          cnt += __popc(mask);
        }
        if (spill) {
          cnt = 0;
#pragma unroll
          for (uint32_t c = 0; c < kBinMaxClasses; ++c) h.mask[c] = 0;
          const uint32_t at = atomicAdd(a.n_spill, 1u);
          a.spill[2ull * at] = (uint32_t)slot;  // slots stay below 2^32 (n_q <= 2^32 - 2 - kTileQ is checked on the host)
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
        for (uint32_t r0 = 0; r0 < total; r0 += kStageIds) {
          uint32_t p = excl;
#pragma unroll
          for (uint32_t c = 0; c < kBinMaxClasses; ++c) {
            uint32_t mask = h.mask[c];
            while (mask) {
              const uint32_t j = __ffs(mask) - 1;
              mask &= mask - 1;
              if (p - r0 < (uint32_t)kStageIds) my_stage[p - r0] = s_id[h.lb[c] + j];
              ++p;
            }
          }
          __syncwarp();
          const uint32_t n_here = min((uint32_t)kStageIds, total - r0);
          for (uint32_t s = lane; s < n_here; s += 32) a.staging[wbase + r0 + s] = my_stage[s];
          __syncwarp();
        }
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
    uint32_t lbs[kBinMaxClasses], ubs[kBinMaxClasses];
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
        if (hit) a.staging[base + done + __popc(bal & ((1u << lane) - 1u))] = a.ids[r];
        done += __popc(bal);
      }
  }
}

// ---------------------------------------------------------------------------------------------------
// Back to query order. Dynamic shared memory: res[kTileQ] uint2 | off[kTileQ + 2] u64 | out[kPlaceCap] u32 |
// loc[kTileQ] u16. Offsets inside a tile are 64-bit: a spilled query may hit every target.
template <bool EMIT>
__global__ void __launch_bounds__(kPlaceThreads, 1) bin_place_kernel(const BinnedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint2* s_res = reinterpret_cast<uint2*>(smem_raw);
  uint64_t* s_off = reinterpret_cast<uint64_t*>(s_res + kTileQ);
  uint32_t* s_out = reinterpret_cast<uint32_t*>(s_off + kTileQ + 2);
  uint16_t* s_loc = reinterpret_cast<uint16_t*>(s_out + kPlaceCap);
  __shared__ uint64_t s_scan[kPlaceThreads / 32 + 1];
  __shared__ uint32_t s_tile;
  __shared__ uint64_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(a.place_ticket, 1u);  // ticket order: every earlier tile has started (look-back)
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t K = a.n_bins;
  const uint16_t* trun = a.trun + (uint64_t)tile * (K + 2);
  const uint32_t n_live = trun[K], n_here = trun[K + 1];  // slots with a bin / queries of the tile
  const uint64_t q0 = (uint64_t)tile * kTileQ;

  for (uint32_t s = tid; s < (uint32_t)kTileQ; s += kPlaceThreads) {
    s_res[s] = s < n_live ? a.sres[q0 + s] : make_uint2(0u, 0u);
    s_loc[s] = s < n_here ? a.bloc[q0 + s] : (uint16_t)0;
    s_off[s] = 0;
  }
  __syncthreads();
  for (uint32_t s = tid; s < n_live; s += kPlaceThreads) s_off[s_loc[s]] = s_res[s].y;  // counts by query position
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
  for (uint64_t w0 = 0; w0 < total; w0 += kPlaceCap) {
    // gather: 32 consecutive slots per warp step; their id lists lie next to each other in the staging area
    for (uint32_t s0 = warp * 32; s0 < n_live; s0 += kPlaceThreads) {
      const uint32_t s = s0 + lane;
      uint2 r = make_uint2(0u, 0u);
      uint64_t dst0 = 0;
      if (s < n_live) { r = s_res[s]; dst0 = s_off[s_loc[s]]; }
      const bool touches = r.y != 0 && dst0 < w0 + kPlaceCap && dst0 + r.y > w0;
      if (!__any_sync(0xffffffffu, touches)) continue;
      const uint32_t incl = warp_incl_scan(r.y, lane);
      const uint32_t m = __shfl_sync(0xffffffffu, incl, 31);
      for (uint32_t fb = 0; fb < m; fb += 32) {
        const uint32_t f = fb + lane;
        const int src = run_owner(incl, f < m ? f : 0u);
        const uint32_t o_incl = __shfl_sync(0xffffffffu, incl, src), o_n = __shfl_sync(0xffffffffu, r.y, src);
        const uint32_t o_beg = __shfl_sync(0xffffffffu, r.x, src);
        const uint64_t o_dst = shfl_u64(dst0, src);
        if (f < m) {
          const uint32_t j = f - (o_incl - o_n);
          const uint64_t dst = o_dst + j - w0;  // wraps far above the window for ids that lie before it
          if (dst < (uint64_t)kPlaceCap) s_out[dst] = o_beg != kNotStored ? a.staging[(uint64_t)o_beg + j] : kNotStored;
        }
      }
    }
    __syncthreads();
    const uint32_t n_out = (uint32_t)min((uint64_t)kPlaceCap, total - w0);
    for (uint32_t p = tid; p < n_out; p += kPlaceThreads) {
      const uint64_t pos = base + w0 + p;
      if (pos < a.capacity) a.hit_target[pos] = s_out[p];
    }
    __syncthreads();
    if (a.hit_query) {  // the query-id column: every query fills its own stretch of the window
      for (uint32_t i = tid; i < n_here; i += kPlaceThreads) {
        const uint64_t b = max(s_off[i], w0), e = min(s_off[i + 1], w0 + (uint64_t)kPlaceCap);
        const uint32_t qid = a.qid_base + (uint32_t)(q0 + i);
        for (uint64_t p = b; p < e; ++p) s_out[p - w0] = qid;
      }
      __syncthreads();
      for (uint32_t p = tid; p < n_out; p += kPlaceThreads) {
        const uint64_t pos = base + w0 + p;
        if (pos < a.capacity) a.hit_query[pos] = s_out[p];
      }
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
  a.low = ix->d_low; a.high = ix->d_high; a.ids = ix->d_id; a.lowhigh = ix->d_lowhigh; a.dir = ix->d_dir;
  a.gtable = ix->d_groups; a.desc = ix->d_bn_desc; a.groups = ix->d_bn_groups; a.cell2bin = ix->d_bn_cell2bin;
  a.n_groups = ix->n_groups; a.n_bins = ix->bn_bins; a.n_cells = ix->bn_cells; a.cell_shift = ix->bn_cell_shift;
  a.max_gval = ix->max_gval; a.n_cls = ix->n_comp;
  a.qgroup = d_qgroup; a.qlow = d_qlow; a.qhigh = d_qhigh; a.n_q = (uint32_t)n_q; a.n_tiles = (uint32_t)n_tiles;
  a.offsets = d_offsets; a.capacity = pair_capacity; a.hit_query = d_hit_query; a.hit_target = d_hit_target;
  a.total = d_total; a.total_mapped = total_mapped; a.qid_base = query_id_base; a.emit = emit;
  a.stage_cap = stage_cap;

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

  const size_t sort_smem = (size_t)kTileQ * 8 + (size_t)((K + 2 + 3) & ~3u) * 4 + (size_t)ix->n_groups * sizeof(BinGroup) +
                           (size_t)kTileQ * 2 + (size_t)((ix->bn_cells + 7) & ~7u) * 2 + (size_t)kBnDirectGroups * 2;
  const size_t probe_smem = (size_t)kBinRowsCap * 12 + (size_t)kProbeWarps * kStageIds * 4 + (size_t)kBinLutCap * 2;
  const size_t place_smem = (size_t)kTileQ * 8 + (size_t)(kTileQ + 2) * 8 + (size_t)kPlaceCap * 4 + (size_t)kTileQ * 2;
  static std::atomic<bool> attrs_set[64];
  if (ix->device < 0 || ix->device >= 64 || !attrs_set[ix->device].load()) {
    BCU_CUDA(cudaFuncSetAttribute(bin_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BCU_CUDA(cudaFuncSetAttribute(bin_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem));
    BCU_CUDA(cudaFuncSetAttribute(bin_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem));
    BCU_CUDA(cudaFuncSetAttribute(bin_place_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
    BCU_CUDA(cudaFuncSetAttribute(bin_place_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
    if (ix->device >= 0 && ix->device < 64) attrs_set[ix->device].store(true);
  }
  if (sort_smem > 200 * 1024) return BCU_NOT_TAKEN;

  const unsigned sort_grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sms * (sort_smem <= 100 * 1024 ? 2 : 1));
  bin_sort_kernel<<<sort_grid, kSortThreads, sort_smem, stream>>>(a);
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
