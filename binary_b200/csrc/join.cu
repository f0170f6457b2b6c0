// K3 + K4 -- the batched overlap join: per-query bound lookup + candidate scan (K3), then
// count -> prefix sum -> scatter of (query_id, target_id) pairs (K4).
//
// Reference being replaced: IntervalTree::find_overlaps / find_overlaps_impl
// (interval_tree.hpp:161-168, 306-334), one recursive pruned walk + vector copies per query, driven
// once per record by sv2nl (mapper.hpp:207-218).
//
// Separate probe and emit launches, no inter-CTA waiting anywhere (a single-pass chained scan was measured
// first: with ~600 resident tiles of random-latency work every tile ends up waiting for the slowest
// in-flight predecessor). The grid is ONE wave (4 CTAs per SM, all resident); K4 and K4b are
// launched as programmatic dependents of their predecessor.
//
//   probe_kernel   K3. The query batch is cut into contiguous chunks, one per CTA; inside a chunk every
//                  WARP takes 128 consecutive queries at a time (4 per lane, one 128-bit load per input
//                  column) and never synchronises with other warps.
//                    bounds: lb = dir[bin(q.low)].lb, ub = dir[bin(q.high)].ub -- ONE 128-bit load when the
//                            query lies inside one directory bin, which also brings the first candidate
//                            row (the directory replaces both binary searches; index_build.cu)
//                    count : rows [lb,ub) are a superset of the hits; the exact predicate
//                            q.low <= t.high && t.low <= q.high (interval_tree.hpp:119-121) is evaluated
//                            on each. Short ranges (<= 31 rows): by the owning lane, its 4 queries
//                            interleaved so their loads overlap, result = a hit bitmask. Long ranges are
//                            listed per chunk and counted in a second phase of the same kernel, one per
//                            lane, by rank arithmetic over a second sorted view of the highs
//                            (count_by_ranks) -- or by a warp scan where a segment has inverted rows, and
//                            always by a scan under a pair filter.
//                  Output: per query 8 bytes of state {lb, hit mask | count} (+ the exact end row of long
//                  ranges), the chunk's list of long ranges, one hit total per CTA.
//   emit_kernel    K4. Same chunks. A CTA first sums the totals of the chunks before its own (its
//                  global base), then walks its chunk 1024 queries at a time: counts from the state ->
//                  warp scan + one barrier -> u64 CSR offsets, in query order; short-range hits are staged
//                  in shared memory at their rank, gathered to target ids, and written to consecutive
//                  addresses.
//   emit_long_kernel K4b. The pairs of the listed long ranges: a warp per range, one row per lane, ballot =
//                  rank, compacted stores (filtered long ranges are emitted by K4 itself).
//
//   direct_kernel  the two-call ABI's second half (offsets supplied by the caller) and the any-overlap
//                  bit: bounds + scan + scatter without any prefix step.
#include "common.cuh"

namespace bcu {

constexpr int kJoinThreads = 256;
constexpr int kJoinWarps = kJoinThreads / 32;
constexpr int kQPT = 4;                            // queries per lane: one 128-bit load per input column
constexpr int kWarpTile = 32 * kQPT;               // 128 queries per warp step
constexpr int kCtaTile = kJoinWarps * kWarpTile;   // 1024 queries per CTA step
#ifndef BCU_PROBE_MB
#define BCU_PROBE_MB 4
#endif
#ifndef BCU_EMIT_MB
#define BCU_EMIT_MB 4
#endif
constexpr int kJoinMinBlocks = 4;
constexpr int kProbeMinBlocks = BCU_PROBE_MB;
constexpr int kEmitMinBlocks = BCU_EMIT_MB;
constexpr uint32_t kScalarMax = 31;                // longer candidate ranges go to the warp-cooperative path
                                                   // (31 = hit-mask bits left beside the flag bit of the state word)
constexpr int kStage = 256;                        // staged hits per warp and round
constexpr int kDirectGroups = 1024;                // group values below this use a direct map
constexpr uint32_t kBigFlag = 0x80000000u;         // state word: bit 31 = long range, low bits = hit count

struct JoinArgs {
  const uint2* __restrict__ lowhigh;
  const uint32_t* __restrict__ high;  // SoA copy of lowhigh[].y, padded: the long-range path streams it
  const uint32_t* __restrict__ ids;
  const DirEntry* __restrict__ dir;
  const uint32_t* __restrict__ hs;    // each segment's `high` values ascending (O(1) long-range count)
  const uint32_t* __restrict__ dirh;  // per directory entry: first index of hs with value >= b*W
  const GroupDesc* __restrict__ groups;
  uint64_t n_rows;  // extents, for the bounds-checked build: rows of the index (arrays are padded by 4),
  uint64_t n_dir;   // directory entries,
  uint64_t n_state; // entries of the probe state arrays / of the long-range list
  uint32_t n_groups;
  uint32_t shifts;  // bin shift of length-class slot c in byte c
  uint32_t max_gval;
  const uint32_t* __restrict__ qgroup;
  const uint32_t* __restrict__ qlow;
  const uint32_t* __restrict__ qhigh;
  uint32_t n_q;    // real queries
  uint32_t n_comp; // length-class slots per query (1, 2 or 4): virtual query v = q * n_comp + slot
  uint32_t n_vq;   // n_q * n_comp virtual queries: what the kernels iterate over
  uint32_t comp_shift;  // log2(n_comp)
  uint32_t chunk;  // queries per CTA chunk (multiple of kCtaTile); direct_kernel: unused
  int vec_ok;      // all query columns (and offsets) are 16-byte aligned
  uint64_t* offsets;
  uint64_t capacity;
  uint32_t* __restrict__ hit_query;
  uint32_t* __restrict__ hit_target;
  uint64_t* total;
  uint8_t* __restrict__ any;
  uint32_t qid_base;
  uint32_t* st_lb;      // [n_q rounded up to kCtaTile] probe state: first candidate row
  uint32_t* st_w;       // [same] probe state: hit bitmask, or kBigFlag | hit count
  uint32_t* st_ub;      // [same] probe state of LONG ranges only: exact end row (untouched otherwise)
  uint64_t* cta_total;  // [gridDim.x] hits per chunk
  uint64_t* total_mapped;  // optional second copy of the total, in mapped pinned host memory
  uint32_t* big_list;   // [padded n_vq] per chunk, from its first slot on: the virtual queries with LONG ranges
  uint32_t* cta_big;    // [gridDim.x] entries of the chunk's list
  uint32_t long_split;  // K4b: CTAs per chunk list
  const uint64_t* base_in;  // optional: offset of the batch's first pair (chunked host pipeline)
  // optional pair filter applied ON TOP of the overlap predicate (sv2nl's check_condition, fused)
  uint32_t filter_kind;     // bcu_filter_kind
  uint32_t filter_diff;
  uint32_t filter_use_strand;
  const uint8_t* __restrict__ qstrand;  // INV: per query, bit0 = strand1 is '+', bit1 = strand2 is '+'
  // stab lists (index_build.cu build_long_lists), lc_seg == nullptr: none
  const uint2* __restrict__ lc_seg;
  const uint32_t* __restrict__ lc_off;
  const uint32_t* __restrict__ lc_row0;
  const uint32_t* __restrict__ lc_high;
  const uint32_t* __restrict__ lc_id;
  uint32_t lc_shift;
  uint64_t lc_entries, lc_bins;
};

struct GroupTables {
  uint32_t g_val[kMaxSmemGroups];  // sorted group values (binary-search mode)
  uint32_t g_nb[kMaxSmemGroups];
  uint64_t g_base[kMaxSmemGroups];
  uint32_t g_re[kMaxSmemGroups];     // row_end of the segment
  uint8_t g_proper[kMaxSmemGroups];  // every row of the segment has low <= high
  uint16_t g_map[kDirectGroups];   // group value -> descriptor index, 0xffff = absent (direct mode)
};

struct StageBuffers {
  uint64_t pos[kJoinWarps][kStage];  // output position of each staged hit relative to the warp's base
                                     // (only used when long ranges sit between the short ones)
  uint2 vq[kJoinWarps][kStage];      // {sorted-row index of the hit, query index inside the warp's 128}
};

// Read-only 128/64-bit loads as ONE instruction. Written as PTX because the compiler otherwise may split
// a struct load whose fields are used under different predicates into dependent scalar loads.
__device__ __forceinline__ uint4 ldg_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  return __shfl_sync(0xffffffffu, (unsigned long long)v, src);
}

__device__ __forceinline__ bool overlaps(uint32_t ql, uint32_t qh, uint32_t tl, uint32_t th) {
  return (ql <= th) & (tl <= qh);
}
__device__ __forceinline__ uint32_t absdiff(uint32_t a, uint32_t b) { return a >= b ? a - b : b - a; }

// The pair predicate. FILT = false: the reference's is_overlap (interval_tree.hpp:119-121), nothing else.
// FILT = true: additionally sv2nl's check_condition for the pair (query = validated NL record, target =
// validated SV record), so that rejected pairs are never counted or written:
//   DUP (mapper.cpp:50-55):  sv contains nl  &&  both ends within diff
//   INV (mapper.cpp:57-79):  neither contains the other  &&  both ends within diff  &&  strand rule
template <bool FILT>
__device__ __forceinline__ bool accept(uint32_t kind, uint32_t diff, uint32_t use_strand, uint32_t strand,
                                       uint32_t ql, uint32_t qh, uint32_t tl, uint32_t th) {
  const bool ov = overlaps(ql, qh, tl, th);
  if (!FILT) return ov;
  const bool t_has_q = (tl <= ql) & (th >= qh);                              // is_contained(sv, nl)
  const bool near = (absdiff(ql, tl) <= diff) & (absdiff(qh, th) <= diff);  // distance_less
  if (kind == BCU_FILTER_SV2NL_DUP) return ov & t_has_q & near;
  const bool q_has_t = (ql <= tl) & (qh >= th);                              // is_contained(nl, sv)
  bool ok = ov & !t_has_q & !q_has_t & near;
  if (use_strand) {
    const bool s1 = strand & 1u, s2 = strand & 2u;
    ok &= (ql <= tl) ? (s1 & !s2) : (!s1 & s2);
  }
  return ok;
}

// group tables -> shared memory, once per CTA (ends with a barrier)
__device__ __forceinline__ void load_group_tables(const JoinArgs& a, GroupTables& tb) {
  const int tid = threadIdx.x;
  const uint32_t n_desc = a.n_groups * a.n_comp;  // table layout: [slot][group]
  const bool in_smem = n_desc <= (uint32_t)kMaxSmemGroups;
  const bool direct = in_smem && a.max_gval < (uint32_t)kDirectGroups;
  if (in_smem) {
    if (direct)
      for (int g = tid; g < kDirectGroups; g += kJoinThreads) tb.g_map[g] = 0xffffu;
    __syncthreads();
    for (uint32_t g = tid; g < n_desc; g += kJoinThreads) {
      const GroupDesc d = a.groups[g];
      tb.g_val[g] = d.gval;
      tb.g_nb[g] = d.nb;
      tb.g_base[g] = d.bin_base;
      tb.g_re[g] = d.row_end;
      tb.g_proper[g] = (uint8_t)d.proper;
      if (direct && g < a.n_groups) tb.g_map[d.gval] = (uint16_t)g;
    }
  }
  __syncthreads();
}

// Loads the 4 consecutive VIRTUAL queries v0..v0+3 of a lane. With one length class a virtual query is
// the query itself (128-bit loads); with n_comp slots, v = q * n_comp + slot, so a lane holds all the
// slots of its 4 / n_comp queries (v0 is a multiple of 4 and n_comp is 1, 2 or 4).
__device__ __forceinline__ void load_queries(const JoinArgs& a, uint32_t v0, uint32_t (&ql)[kQPT],
                                             uint32_t (&qh)[kQPT], uint32_t (&qg)[kQPT]) {
  if (a.n_comp == 1 && a.vec_ok && v0 + kQPT <= a.n_q) {
    uint4 t = *reinterpret_cast<const uint4*>(a.qlow + v0);
    ql[0] = t.x; ql[1] = t.y; ql[2] = t.z; ql[3] = t.w;
    t = *reinterpret_cast<const uint4*>(a.qhigh + v0);
    qh[0] = t.x; qh[1] = t.y; qh[2] = t.z; qh[3] = t.w;
    if (a.qgroup) {
      t = *reinterpret_cast<const uint4*>(a.qgroup + v0);
      qg[0] = t.x; qg[1] = t.y; qg[2] = t.z; qg[3] = t.w;
    } else {
      qg[0] = qg[1] = qg[2] = qg[3] = 0u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      const uint32_t v = v0 + j;
      const uint32_t q = v >> a.comp_shift;
      const bool ok = v >= v0 && q < a.n_q;  // v0 may be past the end (or wrap) when prefetching
      ql[j] = ok ? a.qlow[q] : 0u;
      qh[j] = ok ? a.qhigh[q] : 0u;
      qg[j] = (ok && a.qgroup) ? a.qgroup[q] : 0u;
    }
  }
}

// Bin shift of a length-class slot. Virtual query v = q * n_comp + slot and every lane starts a step at a
// multiple of kQPT >= n_comp, so the slot (hence the shift) of a lane's j-th query is j & (n_comp - 1):
// warp-uniform and loop-invariant, the callers compute the four shifts once.
__device__ __forceinline__ uint32_t slot_shift(const JoinArgs& a, uint32_t slot) {
  return (a.shifts >> (8u * slot)) & 31u;
}

// K3 step 1: candidate row range [lb, lb+len) of one query. `valid` = the query exists.
template <bool FILT>
__device__ __forceinline__ void query_bounds(const JoinArgs& a, const GroupTables& tb, bool valid,
                                             uint32_t slot, uint32_t shift, uint32_t ql, uint32_t qh,
                                             uint32_t qg, uint32_t strand, uint32_t& lb, uint32_t& len,
                                             uint32_t& inline_mask) {  // shift = slot_shift(a, slot)
  const bool in_smem = a.n_groups * a.n_comp <= (uint32_t)kMaxSmemGroups;
  const bool direct = in_smem && a.max_gval < (uint32_t)kDirectGroups;
  lb = 0;
  len = 0;
  inline_mask = 0;
  uint32_t nb = 0;
  uint64_t bin_base = 0;
  if (valid) {
    if (direct) {
      uint32_t gi = qg < (uint32_t)kDirectGroups ? tb.g_map[qg] : 0xffffu;
      if (gi != 0xffffu) { gi += slot * a.n_groups; nb = tb.g_nb[gi]; bin_base = tb.g_base[gi]; }
    } else if (in_smem) {
      uint32_t lo = 0, hi = a.n_groups;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (tb.g_val[mid] < qg) lo = mid + 1; else hi = mid;
      }
      if (lo < a.n_groups && tb.g_val[lo] == qg) {
        lo += slot * a.n_groups;
        nb = tb.g_nb[lo];
        bin_base = tb.g_base[lo];
      }
    } else {
      uint32_t lo = 0, hi = a.n_groups;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a.groups[mid].gval < qg) lo = mid + 1; else hi = mid;
      }
      if (lo < a.n_groups && a.groups[lo].gval == qg) {
        lo += slot * a.n_groups;
        nb = a.groups[lo].nb;
        bin_base = a.groups[lo].bin_base;
      }
    }
  }
  const uint32_t b_lo = ql >> shift;
  if (b_lo < nb) {  // otherwise unknown group, or q.low lies beyond every high of the group
    uint32_t b_hi = qh >> shift;
    if (b_hi >= nb) b_hi = nb - 1u;
    // {lb(b_lo), ub(b_lo + 1), row lb}: all a one-bin query with one candidate needs
    BCU_DEV_ASSERT(bin_base + b_hi < a.n_dir && b_lo < nb);
    const uint4 e = ldg_u4(reinterpret_cast<const uint4*>(a.dir + bin_base + b_lo));
    uint32_t u = e.y;
    if (b_hi != b_lo) u = a.dir[bin_base + b_hi].ub;
    BCU_DEV_ASSERT(e.x <= a.n_rows && u <= a.n_rows);
    lb = e.x;
    len = u > e.x ? u - e.x : 0u;
    inline_mask = (uint32_t)(len > 0 && accept<FILT>(a.filter_kind, a.filter_diff, a.filter_use_strand, strand,
                                                     ql, qh, e.z, e.w));
  }
}

// K3 step 2, short ranges: the 4 queries of a lane advance together (4 independent loads per trip).
template <bool FILT>
__device__ __forceinline__ void scan_short(const JoinArgs& a, const uint32_t (&ql)[kQPT],
                                           const uint32_t (&qh)[kQPT], const uint32_t (&strand)[kQPT],
                                           const uint32_t (&lb)[kQPT], const uint32_t (&len)[kQPT],
                                           uint32_t (&mask)[kQPT]) {
  // on entry mask[j] holds the bit of the row that came inline with the directory entry
  uint32_t max_short = 0;
#pragma unroll
  for (int j = 0; j < kQPT; ++j) {
    if (len[j] > kScalarMax) mask[j] = 0;
    else if (len[j] > max_short) max_short = len[j];
  }
  uint32_t n[kQPT];  // rows of the short range (0 for a long one)
#pragma unroll
  for (int j = 0; j < kQPT; ++j) n[j] = len[j] <= kScalarMax ? len[j] : 0u;
  for (uint32_t k = kInlineRows; k < max_short; ++k) {
    // all four loads are issued before the first is used (written any other way, ptxas under the 64-register
    // cap has been seen to give them one destination register, i.e. to serialise them)
    uint2 t[kQPT];
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      t[j] = make_uint2(0xffffffffu, 0u);
      BCU_DEV_ASSERT(k >= n[j] || (uint64_t)lb[j] + k < a.n_rows);
      if (k < n[j]) t[j] = ldg_u2(a.lowhigh + lb[j] + k);
    }
#pragma unroll
    for (int j = 0; j < kQPT; ++j)
      if (k < n[j])
        mask[j] |= (uint32_t)accept<FILT>(a.filter_kind, a.filter_diff, a.filter_use_strand, strand[j], ql[j],
                                          qh[j], t[j].x, t[j].y) << k;
  }
}

// K3/K4, long ranges. A long range [lb, ub) is first trimmed to the EXACT upper bound by its owning lane
// (the directory's ub is at most about one bin of rows too far; walking back over `low` costs ~1 load), so
// that every remaining row satisfies t.low <= q.high and only t.high >= q.low is left to test. The warp then
// streams the SoA column `high` (4 B per row instead of 8), one row per lane, four coalesced loads in
// flight. When emitting, the matching `id` column is loaded alongside (no dependent gather), a ballot
// ranks the hits and they are written to consecutive slots; the query-id column of a long range is one
// value and is filled with coalesced stores.
struct LongCtx {  // what the long-range path needs, passed BY VALUE into the out-of-line function so
                  // that the callers' per-query arrays stay in registers on the (common) short path
  const uint32_t* high;
  const uint32_t* ids;
  const uint2* lowhigh;  // FILT only: the filter needs t.low as well
  uint32_t* hit_target;
  uint32_t* hit_query;
  uint64_t capacity;
  uint32_t qid_base;
  uint32_t comp_shift;
  uint32_t filter_kind, filter_diff, filter_use_strand;
  uint64_t n_rows;
};

struct LongRange {
  uint32_t lb, ub, ql;  // rows [lb, ub), query low
  uint32_t qh, strand;  // FILT only
  uint32_t cnt;         // EMIT: number of hits (from the probe state)
  uint32_t qid;
  uint64_t base;        // EMIT: output position of the first hit
};

__device__ __forceinline__ uint32_t exact_upper_bound(const JoinArgs& a, uint32_t lb, uint32_t len,
                                                      uint32_t qh) {
  uint32_t u = lb + len;
  BCU_DEV_ASSERT((uint64_t)lb + len <= a.n_rows);
  while (u > lb && a.lowhigh[u - 1].x > qh) --u;
  return u;
}

// O(1) hit count of a long range by rank arithmetic. When every row of the segment is proper (low <= high)
// and so is the query, the rows that end before the query, {high < q.low}, are a subset of the rows that
// start in time, {low <= q.high}; hence
//     hits = #{low <= q.high} - #{high < q.low} = (ub_exact - row_begin) - (i - row_begin) = ub_exact - i,
// with i = lower_bound of q.low in the segment's ascending `high` view (directory dirh + a short walk).
// Returns kNoRank when the shortcut does not apply (inverted rows or query): the caller scans instead.
constexpr uint32_t kNoRank = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t count_by_ranks(const JoinArgs& a, const GroupTables& tb, uint32_t slot,
                                                   uint32_t ql, uint32_t qh, uint32_t qg, uint32_t ub_exact) {
  if (ql > qh) return kNoRank;
  const bool in_smem = a.n_groups * a.n_comp <= (uint32_t)kMaxSmemGroups;
  const bool direct = in_smem && a.max_gval < (uint32_t)kDirectGroups;
  uint32_t gi;
  if (direct) {
    gi = tb.g_map[qg];  // the probe found candidates, so the group exists and qg is in range
  } else {
    uint32_t lo = 0, hi = a.n_groups;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      const uint32_t v = in_smem ? tb.g_val[mid] : a.groups[mid].gval;
      if (v < qg) lo = mid + 1; else hi = mid;
    }
    gi = lo;
  }
  gi += slot * a.n_groups;
  const bool proper = in_smem ? tb.g_proper[gi] != 0 : a.groups[gi].proper != 0;
  if (!proper) return kNoRank;
  const uint64_t bin_base = in_smem ? tb.g_base[gi] : a.groups[gi].bin_base;
  const uint32_t row_end = in_smem ? tb.g_re[gi] : a.groups[gi].row_end;
  BCU_DEV_ASSERT(bin_base + (ql >> slot_shift(a, slot)) < a.n_dir && row_end <= a.n_rows);
  uint32_t i = a.dirh[bin_base + (ql >> slot_shift(a, slot))];  // b_lo < nb: the range is non-empty
  while (i < row_end && a.hs[i] < ql) ++i;
  return ub_exact > i ? ub_exact - i : 0u;
}

// bits k = 0..3: row r+k lies inside [lb, ub)
__device__ __forceinline__ uint32_t rows_in_range(uint32_t r, uint32_t lb, uint32_t ub) {
  const uint32_t n_hi = ub > r ? min(ub - r, 4u) : 0u;
  const uint32_t n_lo = lb > r ? min(lb - r, 4u) : 0u;
  return ((1u << n_hi) - 1u) & ~((1u << n_lo) - 1u);
}
__device__ __forceinline__ uint32_t reaches(uint32_t ql, const uint4& h) {
  return (uint32_t)(ql <= h.x) | ((uint32_t)(ql <= h.y) << 1) | ((uint32_t)(ql <= h.z) << 2) |
         ((uint32_t)(ql <= h.w) << 3);
}

// COUNT only, no filter: 4 adjacent rows per lane per 128-bit load of `high`, TWO loads in flight (256 rows per
// trip); a lane-local popcount is all that is needed, the warp reduces once per range. (For counting this
// beats the one-row-per-lane loop below; for emitting it is the other way round, measured.)
__device__ __forceinline__ uint32_t count_long(const LongCtx& c, int lane, const LongRange& R) {
  const uint4* high4 = reinterpret_cast<const uint4*>(c.high);
  uint32_t count = 0;
  for (uint32_t r = (R.lb & ~3u) + 4u * lane; (r - 4u * lane) < R.ub; r += 256) {
    const uint32_t r2 = r + 128;
    uint4 h0 = make_uint4(0, 0, 0, 0), h1 = make_uint4(0, 0, 0, 0);
    BCU_DEV_ASSERT(R.ub <= c.n_rows);  // a 128-bit load may reach up to 3 rows past ub: the arrays are padded by 4
    if (r < R.ub) h0 = ldg_u4(high4 + (r >> 2));
    if (r2 < R.ub) h1 = ldg_u4(high4 + (r2 >> 2));
    count += __popc(rows_in_range(r, R.lb, R.ub) & reaches(R.ql, h0));
    count += __popc(rows_in_range(r2, R.lb, R.ub) & reaches(R.ql, h1));
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
  return count;
}

// One long range, scanned by the whole warp: ONE row per lane and trip (the ballot then IS the rank, no
// shuffle scan), four trips (128 rows) loaded before the first is consumed. Measured on the dense config: fewer instructions per hit than a 4-rows-per-lane / 128-bit-load variant, whose mask
// building, shuffle scan and 8 predicated stores per trip outweighed the saved load instructions (the
// kernel is issue- and latency-bound, not L2-bound).
// FILT: rows are read as {low, high} (8 B per lane) and go through accept<true>.
constexpr int kTrips = 4;
struct LongRegs { uint32_t hi[kTrips], lo[kTrips], id[kTrips]; };

template <bool EMIT, bool FILT>
__device__ __forceinline__ void long_preload(const LongCtx& c, int lane, const LongRange& R, uint32_t r0,
                                             LongRegs& g) {
#pragma unroll
  for (int t = 0; t < kTrips; ++t) {
    const uint32_t r = r0 + 32 * t + lane;
    g.hi[t] = 0;
    g.lo[t] = 0xffffffffu;
    g.id[t] = 0;
    BCU_DEV_ASSERT(R.ub <= c.n_rows);
    if (r < R.ub) {
      if (FILT) {
        const uint2 v = ldg_u2(c.lowhigh + r);
        g.lo[t] = v.x;
        g.hi[t] = v.y;
      } else {
        g.hi[t] = __ldg(c.high + r);
      }
      if (EMIT) g.id[t] = __ldg(c.ids + r);
    }
  }
}

// count: COUNT = lane-local hit count; EMIT = hits of the range written so far (warp-uniform)
template <bool EMIT, bool FILT>
__device__ __forceinline__ void long_consume(const LongCtx& c, int lane, const LongRange& R, uint32_t r0,
                                             const LongRegs& g, uint32_t* out, uint32_t lim, uint32_t& count) {
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int t = 0; t < kTrips; ++t) {
    if (t > 0 && r0 + 32 * t >= R.ub) break;  // warp-uniform: do not pay for trips past the end of the range
    const uint32_t r = r0 + 32 * t + lane;
    bool hit = r < R.ub;
    if (FILT) hit = hit && accept<true>(c.filter_kind, c.filter_diff, c.filter_use_strand, R.strand, R.ql, R.qh,
                                        g.lo[t], g.hi[t]);
    else hit = hit && R.ql <= g.hi[t];
    if (!EMIT) {
      count += hit;
    } else {
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      const uint32_t p = count + __popc(bal & lt);
      BCU_DEV_ASSERT(!(hit && p < lim) || R.base + p < c.capacity);
      if (hit && p < lim) out[p] = g.id[t];
      count += __popc(bal);
    }
  }
}

// All long ranges of the warp's 128 queries. Lane-local inputs per query j (packed in
// vectors, by value): bit j of `bigbits`, the exact row range [lb.j, ub.j), ql.j; EMIT also needs cnt.j
// (from the probe) and the absolute output position pos.j. COUNT returns cnt.j of the long ranges.
template <bool EMIT, bool FILT>
__device__ __noinline__ uint4 long_ranges(LongCtx c, int lane, uint32_t bigbits, uint4 lb4, uint4 ub4,
                                          uint4 ql4, uint4 cnt4, uint64_t pos_0, uint64_t pos_1,
                                          uint64_t pos_2, uint64_t pos_3, uint32_t vq0, uint4 qh4,
                                          uint32_t strand4) {  // strand4: byte j = strand bits of query j
  const uint32_t lb[kQPT] = {lb4.x, lb4.y, lb4.z, lb4.w};
  const uint32_t ub[kQPT] = {ub4.x, ub4.y, ub4.z, ub4.w};
  const uint32_t ql[kQPT] = {ql4.x, ql4.y, ql4.z, ql4.w};
  const uint32_t qh[kQPT] = {qh4.x, qh4.y, qh4.z, qh4.w};
  uint32_t cnt[kQPT] = {cnt4.x, cnt4.y, cnt4.z, cnt4.w};
  const uint64_t pos0[kQPT] = {pos_0, pos_1, pos_2, pos_3};
  if (!EMIT && !FILT) {  // plain counting: 128-bit loads, one range after the other
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      unsigned todo = __ballot_sync(0xffffffffu, (bigbits >> j) & 1u);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        LongRange R;
        R.lb = __shfl_sync(0xffffffffu, lb[j], src);
        R.ub = __shfl_sync(0xffffffffu, ub[j], src);
        R.ql = __shfl_sync(0xffffffffu, ql[j], src);
        const uint32_t n = count_long(c, lane, R);
        if (lane == src) cnt[j] = n;
      }
    }
    return make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
  }

  // ---- emit / filtered count: one range after the other. (Software-pipelining the ranges one ahead was
  // measured and lost: its bookkeeping costs more issue slots than the latency it hides.)
#pragma unroll
  for (int j = 0; j < kQPT; ++j) {
    unsigned todo = __ballot_sync(0xffffffffu, (bigbits >> j) & 1u);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      LongRange R;
      R.lb = __shfl_sync(0xffffffffu, lb[j], src);
      R.ub = __shfl_sync(0xffffffffu, ub[j], src);
      R.ql = __shfl_sync(0xffffffffu, ql[j], src);
      R.qh = R.strand = 0;
      if (FILT) {
        R.qh = __shfl_sync(0xffffffffu, qh[j], src);
        R.strand = (__shfl_sync(0xffffffffu, strand4, src) >> (8 * j)) & 0xffu;
      }
      R.qid = c.qid_base + ((vq0 + (uint32_t)src * kQPT + j) >> c.comp_shift);  // vq0 = warp's first virtual query
      R.cnt = 0;
      R.base = 0;
      if (EMIT) {
        R.cnt = __shfl_sync(0xffffffffu, cnt[j], src);
        R.base = shfl_u64(pos0[j], src);
      }
      uint32_t* out = c.hit_target + R.base;  // 32-bit ranks against a per-range pointer
      const uint32_t lim = c.capacity > R.base ? (uint32_t)min(c.capacity - R.base, (uint64_t)0xffffffffu) : 0u;
      uint32_t count = 0;
      for (uint32_t r0 = R.lb; r0 < R.ub; r0 += 32 * kTrips) {
        LongRegs g;
        long_preload<EMIT, FILT>(c, lane, R, r0, g);
        long_consume<EMIT, FILT>(c, lane, R, r0, g, out, lim, count);
      }
      if (!EMIT) {
#pragma unroll
        for (int off = 16; off; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
        if (lane == src) cnt[j] = count;
      } else if (c.hit_query) {  // the query-id column of a range is one value: coalesced fill
        // (fusing this store into the per-hit store above was measured: slower in the emit kernel)
        uint32_t* q = c.hit_query + R.base;
        const uint32_t n = min(R.cnt, lim);
        for (uint32_t k = lane; k < n; k += 32) q[k] = R.qid;
      }
    }
  }
  return make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
}

__device__ __forceinline__ LongCtx long_ctx(const JoinArgs& a) {
  LongCtx c;
  c.high = a.high;
  c.ids = a.ids;
  c.hit_target = a.hit_target;
  c.hit_query = a.hit_query;
  c.lowhigh = a.lowhigh;
  c.capacity = a.capacity;
  c.qid_base = a.qid_base;
  c.comp_shift = a.comp_shift;
  c.filter_kind = a.filter_kind;
  c.filter_diff = a.filter_diff;
  c.filter_use_strand = a.filter_use_strand;
  c.n_rows = a.n_rows;
  return c;
}
__device__ __forceinline__ uint32_t pack_bits(const bool (&b)[kQPT]) {
  return (uint32_t)b[0] | ((uint32_t)b[1] << 1) | ((uint32_t)b[2] << 2) | ((uint32_t)b[3] << 3);
}

// K4 scatter, short ranges of one warp (128 queries): stage each hit's row at its rank among the warp's
// short-range hits, then gather rows -> target ids with independent loads and write to consecutive
// addresses. off[j] = output position of query j's first hit, relative to `base`.
// GAPS = false: the warp has no long range, so a hit's output position IS its rank (no position array).
template <bool GAPS>
__device__ __forceinline__ void emit_short(const JoinArgs& a, StageBuffers& st, int warp, int lane,
                                           const uint32_t (&mask)[kQPT], const uint32_t (&lb)[kQPT],
                                           const uint64_t (&off)[kQPT], uint64_t base, uint32_t vq0) {
  const uint32_t lane_hits = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]);
  uint32_t incl = lane_hits;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const uint32_t staged_total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t first_slot = incl - lane_hits;
  for (uint32_t r0 = 0; r0 < staged_total; r0 += kStage) {
    uint32_t slot = first_slot;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      uint32_t m = mask[j];
      const uint32_t n_j = __popc(m);
      if (n_j && slot + n_j > r0 && slot < r0 + kStage) {  // this query has hits inside the round's window
        const uint32_t qi = (uint32_t)(lane * kQPT + j);
        uint32_t s = slot - r0;  // wraps below the window: caught by the unsigned compare
        uint64_t pos = off[j];
        // most queries have exactly one hit: handle it without the loop
        if (s < (uint32_t)kStage) {
          st.vq[warp][s] = make_uint2(lb[j] + (__ffs(m) - 1), qi);
          if (GAPS) st.pos[warp][s] = pos;
        }
        m &= m - 1;
        while (m) {
          ++s;
          ++pos;
          if (s < (uint32_t)kStage) {
            st.vq[warp][s] = make_uint2(lb[j] + (__ffs(m) - 1), qi);
            if (GAPS) st.pos[warp][s] = pos;
          }
          m &= m - 1;
        }
      }
      slot += n_j;
    }
    __syncwarp();
    const uint32_t n_here = min((uint32_t)kStage, staged_total - r0);
    // the id gather id[row] is the one random access of this kernel: four per lane are issued before the
    // first is stored (a plain loop serialised them: 24 % of the kernel's stall samples sat on this line)
    for (uint32_t s0 = 0; s0 < n_here; s0 += 128) {
      uint2 v[4];
      uint32_t id[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t s = s0 + 32 * t + lane;
        v[t] = make_uint2(0, 0);
        id[t] = 0;
        if (s < n_here) {
          v[t] = st.vq[warp][s];
          BCU_DEV_ASSERT(v[t].x < a.n_rows);
          id[t] = __ldg(a.ids + v[t].x);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t s = s0 + 32 * t + lane;
        if (s < n_here) {
          const uint64_t pos = base + (GAPS ? st.pos[warp][s] : (uint64_t)(r0 + s));
          if (pos < a.capacity) {
            a.hit_target[pos] = id[t];
            if (a.hit_query) a.hit_query[pos] = a.qid_base + ((vq0 + v[t].y) >> a.comp_shift);
          }
        }
      }
    }
    __syncwarp();
  }
}

// Programmatic dependent launch (sm_90+): the producer lets the next kernel of the stream be placed early,
// the consumer waits here until the producer grid has completed and its writes are visible.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// ---------------------------------------------------------------------------------------------------
// K3: probe. No barriers inside the chunk loop. A long range (> kScalarMax rows) only leaves its marker and
// inexact end in the state and its query number in the chunk's list, for phase 2 below and for K4b; under a
// filter (which must see every pair) the warp scans it in place instead.
template <bool FILT>
__global__ void __launch_bounds__(kJoinThreads, kProbeMinBlocks) probe_kernel(const JoinArgs a) {
  __shared__ GroupTables tb;
  __shared__ uint64_t s_warp_total[kJoinWarps];
  __shared__ uint32_t s_big;  // long ranges listed by this CTA so far
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_big = 0;
  grid_launch_dependents();
  load_group_tables(a, tb);  // ends with a barrier

  const uint64_t chunk_begin = (uint64_t)blockIdx.x * a.chunk;
  const uint64_t chunk_end = min(chunk_begin + (uint64_t)a.chunk, (uint64_t)a.n_vq);
  uint32_t* const list = a.big_list + chunk_begin;
  uint64_t acc = 0;  // hits of this lane's queries over the whole chunk (short ranges)
  const uint32_t shift_of[kQPT] = {slot_shift(a, 0), slot_shift(a, 1u & (a.n_comp - 1u)),
                                   slot_shift(a, 2u & (a.n_comp - 1u)), slot_shift(a, 3u & (a.n_comp - 1u))};

  uint64_t w0 = chunk_begin + (uint64_t)warp * kWarpTile;
  uint32_t nql[kQPT], nqh[kQPT], nqg[kQPT];  // prefetched queries of the next step
  if (w0 < chunk_end) load_queries(a, (uint32_t)w0 + lane * kQPT, nql, nqh, nqg);
  for (; w0 < chunk_end; w0 += kCtaTile) {
    const uint32_t q0 = (uint32_t)w0 + (uint32_t)lane * kQPT;
    uint32_t ql[kQPT], qh[kQPT], qg[kQPT];
#pragma unroll
    for (int j = 0; j < kQPT; ++j) { ql[j] = nql[j]; qh[j] = nqh[j]; qg[j] = nqg[j]; }
    if (w0 + kCtaTile < chunk_end) load_queries(a, q0 + kCtaTile, nql, nqh, nqg);

    uint32_t lb[kQPT], len[kQPT], w[kQPT], strand[kQPT];
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      strand[j] = 0;
      if (FILT && a.qstrand && q0 + j < a.n_vq) strand[j] = a.qstrand[(q0 + j) >> a.comp_shift];
      query_bounds<FILT>(a, tb, q0 + j < a.n_vq, j & (a.n_comp - 1u), shift_of[j], ql[j], qh[j], qg[j], strand[j], lb[j],
                         len[j], w[j]);
    }
    scan_short<FILT>(a, ql, qh, strand, lb, len, w);
    bool big[kQPT];
    uint32_t ub[kQPT];
    bool lane_big = false;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      acc += __popc(w[j]);
      big[j] = len[j] > kScalarMax;
      lane_big |= big[j];
      ub[j] = lb[j] + len[j];
    }
    if (__any_sync(0xffffffffu, lane_big)) {
      if (!FILT) {
        // list them (warp scan of the per-lane counts, one shared-memory atomic per warp and step)
        const uint32_t n_big = (uint32_t)big[0] + big[1] + big[2] + big[3];
        uint32_t incl = n_big;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += y;
        }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(&s_big, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - n_big;
#pragma unroll
        for (int j = 0; j < kQPT; ++j)
          if (big[j]) {
            BCU_DEV_ASSERT(chunk_begin + base < a.n_state && q0 + j < a.n_state);
            list[base++] = q0 + j;
            a.st_ub[q0 + j] = ub[j];
            w[j] = kBigFlag;
          }
      } else {
        // filtered: the whole warp scans them here, one after the other
#pragma unroll
        for (int j = 0; j < kQPT; ++j)
          if (big[j]) ub[j] = exact_upper_bound(a, lb[j], len[j], qh[j]);
        const uint4 c4 = long_ranges<false, FILT>(
            long_ctx(a), lane, pack_bits(big), make_uint4(lb[0], lb[1], lb[2], lb[3]),
            make_uint4(ub[0], ub[1], ub[2], ub[3]), make_uint4(ql[0], ql[1], ql[2], ql[3]), make_uint4(0, 0, 0, 0),
            0, 0, 0, 0, 0, make_uint4(qh[0], qh[1], qh[2], qh[3]),
            strand[0] | (strand[1] << 8) | (strand[2] << 16) | (strand[3] << 24));
        const uint32_t cl[kQPT] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int j = 0; j < kQPT; ++j)
          if (big[j]) { acc += cl[j]; w[j] = kBigFlag | cl[j]; a.st_ub[q0 + j] = ub[j]; }
      }
    }
    // state: 8 bytes per query, 128-bit stores (the arrays are padded to a multiple of kCtaTile)
    BCU_DEV_ASSERT((uint64_t)q0 + kQPT <= a.n_state);
    *reinterpret_cast<uint4*>(a.st_lb + q0) = make_uint4(lb[0], lb[1], lb[2], lb[3]);
    *reinterpret_cast<uint4*>(a.st_w + q0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  // Phase 2: the long ranges this CTA listed, one per LANE: exact end of the candidate rows, then the hit
  // count by rank arithmetic (count_by_ranks); the ranges whose segment does not allow it are scanned by
  // the whole warp. The other CTAs of the SM are still in their phase 1, which hides these dependent loads.
  if (!FILT) {
    __syncthreads();
    const uint32_t n_list = s_big;
    for (uint32_t i0 = warp * 32; i0 < n_list; i0 += kJoinThreads) {  // warp-uniform trip count
      const bool have = i0 + lane < n_list;
      uint32_t v = 0, lb = 0, ub = 0, ql = 0, qh = 0, cnt = 0;
      bool scan = false;
      if (have) {
        v = list[i0 + lane];
        const uint32_t q = v >> a.comp_shift;
        lb = a.st_lb[v];
        ql = a.qlow[q];
        qh = a.qhigh[q];
        ub = exact_upper_bound(a, lb, a.st_ub[v] - lb, qh);
        a.st_ub[v] = ub;
        cnt = count_by_ranks(a, tb, v & (a.n_comp - 1u), ql, qh, a.qgroup ? a.qgroup[q] : 0u, ub);
        scan = cnt == kNoRank;
      }
      if (__any_sync(0xffffffffu, scan)) {
        const uint4 c4 = long_ranges<false, false>(long_ctx(a), lane, scan ? 1u : 0u, make_uint4(lb, 0, 0, 0),
                                                   make_uint4(ub, 0, 0, 0), make_uint4(ql, 0, 0, 0),
                                                   make_uint4(0, 0, 0, 0), 0, 0, 0, 0, 0, make_uint4(qh, 0, 0, 0), 0);
        if (scan) cnt = c4.x;
      }
      if (have) {
        a.st_w[v] = kBigFlag | cnt;
        acc += cnt;
      }
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += shfl_u64(acc, lane ^ off);
  if (lane == 0) s_warp_total[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    uint64_t t = 0;
#pragma unroll
    for (int w = 0; w < kJoinWarps; ++w) t += s_warp_total[w];
    a.cta_total[blockIdx.x] = t;
    a.cta_big[blockIdx.x] = FILT ? 0u : s_big;
  }
}

// ---------------------------------------------------------------------------------------------------
// K4: prefix sum + scatter. One barrier per 1024 queries, no waiting on other CTAs.
template <bool EMIT, bool FILT>
__global__ void __launch_bounds__(kJoinThreads, kEmitMinBlocks) emit_kernel(const JoinArgs a) {
  __shared__ StageBuffers st;
  __shared__ uint64_t s_warp_total[2][kJoinWarps];
  __shared__ uint64_t s_red[kJoinWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  grid_dependency_wait();    // K3's state and totals
  grid_launch_dependents();  // K4b may be placed behind this grid

  // global base of this chunk = hits of all chunks before it
  uint64_t part = 0;
  for (uint32_t c = tid; c < blockIdx.x; c += kJoinThreads) part += a.cta_total[c];
#pragma unroll
  for (int off = 16; off; off >>= 1) part += shfl_u64(part, lane ^ off);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  uint64_t running = a.base_in ? *a.base_in : 0ull;
#pragma unroll
  for (int w = 0; w < kJoinWarps; ++w) running += s_red[w];

  const uint64_t chunk_begin = (uint64_t)blockIdx.x * a.chunk;
  const uint64_t chunk_end = min(chunk_begin + (uint64_t)a.chunk, (uint64_t)a.n_vq);
  if (blockIdx.x == gridDim.x - 1 && tid == 0) {  // grand total
    const uint64_t t = running + a.cta_total[blockIdx.x];
    a.offsets[a.n_q] = t;
    if (a.total) *a.total = t;
    if (a.total_mapped) *a.total_mapped = t;  // pinned host memory: visible to the host once the grid is done
  }

  uint4 n_lb = make_uint4(0, 0, 0, 0), n_w = make_uint4(0, 0, 0, 0);  // prefetched state of the next step
  {
    const uint64_t q = chunk_begin + (uint64_t)tid * kQPT;
    if (q < chunk_end) {
      n_lb = *reinterpret_cast<const uint4*>(a.st_lb + q);
      n_w = *reinterpret_cast<const uint4*>(a.st_w + q);
    }
  }
  int par = 0;
  bool seen_big = false;
  for (uint64_t t0 = chunk_begin; t0 < chunk_end; t0 += kCtaTile, par ^= 1) {
    const uint32_t w0 = (uint32_t)t0 + (uint32_t)warp * kWarpTile;
    const uint32_t q0 = w0 + (uint32_t)lane * kQPT;
    const uint32_t lb[kQPT] = {n_lb.x, n_lb.y, n_lb.z, n_lb.w};
    const uint32_t w[kQPT] = {n_w.x, n_w.y, n_w.z, n_w.w};
    if (t0 + kCtaTile + (uint64_t)tid * kQPT < chunk_end) {
      n_lb = *reinterpret_cast<const uint4*>(a.st_lb + q0 + kCtaTile);
      n_w = *reinterpret_cast<const uint4*>(a.st_w + q0 + kCtaTile);
    } else {
      n_lb = make_uint4(0, 0, 0, 0);
      n_w = make_uint4(0, 0, 0, 0);
    }
    uint32_t cnt[kQPT], mask[kQPT];
    bool lane_big = false;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      const bool big = (w[j] & kBigFlag) != 0;
      lane_big |= big;
      mask[j] = big ? 0u : w[j];
      cnt[j] = big ? (w[j] & ~kBigFlag) : (uint32_t)__popc(w[j]);
    }
    // ---- prefix: lane -> warp (shuffles) -> CTA step (shared memory, one barrier) -> running base ----
    const uint64_t lane_sum = (uint64_t)cnt[0] + cnt[1] + cnt[2] + cnt[3];
    uint64_t incl = lane_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t v = __shfl_up_sync(0xffffffffu, (unsigned long long)incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_total[par][warp] = incl;
    __syncthreads();
    uint64_t warp_base = running;
#pragma unroll
    for (int ww = 0; ww < kJoinWarps; ++ww) {
      const uint64_t t = s_warp_total[par][ww];
      if (ww < warp) warp_base += t;
      running += t;
    }
    uint64_t off[kQPT];  // relative to warp_base
    {
      uint64_t run = incl - lane_sum;
#pragma unroll
      for (int j = 0; j < kQPT; ++j) { off[j] = run; run += cnt[j]; }
    }
    if (a.n_comp == 1 && a.vec_ok && q0 + kQPT <= a.n_q) {
      ulonglong2* o = reinterpret_cast<ulonglong2*>(a.offsets + q0);
      o[0] = make_ulonglong2(warp_base + off[0], warp_base + off[1]);
      o[1] = make_ulonglong2(warp_base + off[2], warp_base + off[3]);
    } else {  // the CSR offset of a query is the position of its FIRST slot's first hit
#pragma unroll
      for (int j = 0; j < kQPT; ++j) {
        const uint32_t v = q0 + j;
        if ((v & (a.n_comp - 1u)) == 0 && v < a.n_vq) a.offsets[v >> a.comp_shift] = warp_base + off[j];
      }
    }
    if (!EMIT) continue;

    // ---- scatter ------------------------------------------------------------------------------------
    const bool warp_big = __any_sync(0xffffffffu, lane_big);
    if (warp_big) emit_short<true>(a, st, warp, lane, mask, lb, off, warp_base, w0);
    else emit_short<false>(a, st, warp, lane, mask, lb, off, warp_base, w0);
    seen_big |= warp_big;
  }
  if (!EMIT || !seen_big) return;

  // ---- long ranges, in a second walk over the warp's steps ---------------------------------------------
  // Kept out of the loop above so that its register needs (and the out-of-line call) do not spill into
  // the hot short-range path. Everything comes from the probe state, the query column and the offsets
  // this very thread wrote a moment ago.
  if (!FILT) return;  // without a filter the long ranges are listed: K4b emits them
  for (uint64_t t0 = chunk_begin; t0 < chunk_end; t0 += kCtaTile) {
    const uint32_t w0 = (uint32_t)t0 + (uint32_t)warp * kWarpTile;
    const uint32_t q0 = w0 + (uint32_t)lane * kQPT;
    uint4 w4 = make_uint4(0, 0, 0, 0);
    if (t0 + (uint64_t)tid * kQPT < chunk_end) w4 = *reinterpret_cast<const uint4*>(a.st_w + q0);
    const uint32_t w[kQPT] = {w4.x, w4.y, w4.z, w4.w};
    if (!__any_sync(0xffffffffu, ((w4.x | w4.y | w4.z | w4.w) & kBigFlag) != 0)) continue;
    uint32_t ql[kQPT], qh[kQPT], blb[kQPT], bub[kQPT], cnt[kQPT];
    uint64_t pos[kQPT];
    bool big[kQPT];
    uint32_t strand4 = 0;
    uint64_t run = 0;  // hits of the earlier slots of the same query (all slots of a query sit in one lane)
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      const uint32_t v = q0 + j;
      big[j] = (w[j] & kBigFlag) != 0;
      cnt[j] = big[j] ? (w[j] & ~kBigFlag) : (uint32_t)__popc(w[j]);
      if ((v & (a.n_comp - 1u)) == 0) run = 0;
      ql[j] = big[j] ? a.qlow[v >> a.comp_shift] : 0u;
      qh[j] = (FILT && big[j]) ? a.qhigh[v >> a.comp_shift] : 0u;
      if (FILT && big[j] && a.qstrand) strand4 |= (uint32_t)a.qstrand[v >> a.comp_shift] << (8 * j);
      blb[j] = big[j] ? a.st_lb[v] : 0u;
      bub[j] = big[j] ? a.st_ub[v] : 0u;
      pos[j] = big[j] ? a.offsets[v >> a.comp_shift] + run : 0ull;
      run += cnt[j];
    }
    long_ranges<true, FILT>(long_ctx(a), lane, pack_bits(big), make_uint4(blb[0], blb[1], blb[2], blb[3]),
                            make_uint4(bub[0], bub[1], bub[2], bub[3]), make_uint4(ql[0], ql[1], ql[2], ql[3]),
                            make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]), pos[0], pos[1], pos[2], pos[3], w0,
                            make_uint4(qh[0], qh[1], qh[2], qh[3]), strand4);
  }
}

// ---------------------------------------------------------------------------------------------------
// K4b: the pairs of the listed long ranges (no filter). Same grid and chunks as K3: a CTA walks its own
// chunk's list, 32 entries per warp and round (one per lane), then the warp streams the ranges one after
// the other (long_preload / long_consume). A kernel of its own because this part is bound by the latency
// of one range after the other per warp: it needs far fewer registers than K4, so twice the warps per SM.
#ifndef BCU_LONG_MB
#define BCU_LONG_MB 6
#endif
constexpr int kLongMinBlocks = BCU_LONG_MB;
__global__ void __launch_bounds__(kJoinThreads, kLongMinBlocks) emit_long_kernel(const JoinArgs a) {
  grid_dependency_wait();  // K4's offsets (and through it K3's lists)
  // a.long_split CTAs share one chunk's list (this kernel runs more CTAs per SM than K3, whose grid made the lists)
  const uint32_t chunk_id = blockIdx.x / a.long_split, part = blockIdx.x % a.long_split;
  const uint32_t n_list = a.cta_big[chunk_id];
  if (n_list == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* const list = a.big_list + (uint64_t)chunk_id * a.chunk;
  const LongCtx c = long_ctx(a);
  for (uint32_t i0 = (part * kJoinWarps + warp) * 32; i0 < n_list; i0 += kJoinThreads * a.long_split) {  // warp-uniform
    const bool have = i0 + lane < n_list;
    LongRange mine;
    mine.lb = mine.ub = mine.ql = mine.qh = mine.strand = mine.cnt = mine.qid = 0;
    mine.base = 0;
    uint32_t cov_b = 0, cov_e = 0;  // the block of the bin holding q.low, when the index has stab lists
    if (have) {
      const uint32_t v = list[i0 + lane], q = v >> a.comp_shift;
      mine.lb = a.st_lb[v];
      mine.ub = a.st_ub[v];
      mine.ql = a.qlow[q];
      mine.cnt = a.st_w[v] & ~kBigFlag;
      mine.qid = a.qid_base + q;
      uint64_t pos = a.offsets[q];  // + the hits of the query's earlier length-class slots
      for (uint32_t u = v & ~(a.n_comp - 1u); u < v; ++u) {
        const uint32_t w = a.st_w[u];
        pos += (w & kBigFlag) ? (w & ~kBigFlag) : (uint32_t)__popc(w);
      }
      mine.base = pos;
      if (a.lc_seg && mine.cnt != 0 && mine.ql <= a.qhigh[q]) {  // (an inverted query keeps the plain scan)
        const uint32_t qg = a.qgroup ? a.qgroup[q] : 0u;
        uint32_t lo = 0, hi = a.n_groups;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (a.groups[mid].gval < qg) lo = mid + 1; else hi = mid;
        }
        const uint2 seg = lo < a.n_groups ? a.lc_seg[(v & (a.n_comp - 1u)) * a.n_groups + lo] : make_uint2(0u, 0u);
        const uint32_t b = mine.ql >> a.lc_shift;
        if (b + 1 < seg.y) {  // (seg.y counts the empty end bin)
          BCU_DEV_ASSERT((uint64_t)seg.x + b + 1 < a.lc_bins);
          // the bin's block = stab list (the hits that start before the bin) + the bin's own rows; the rows of the
          // following bins up to ub, if any, come from the plain columns
          const uint32_t r1 = a.lc_row0[seg.x + b + 1];
          cov_b = a.lc_off[seg.x + b];
          cov_e = a.lc_off[seg.x + b + 1] - (r1 > mine.ub ? r1 - mine.ub : 0u);
          mine.lb = max(mine.lb, r1);
        }
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, have && mine.cnt != 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      LongRange R;
      R.lb = __shfl_sync(0xffffffffu, mine.lb, src);
      R.ub = __shfl_sync(0xffffffffu, mine.ub, src);
      R.ql = __shfl_sync(0xffffffffu, mine.ql, src);
      R.cnt = __shfl_sync(0xffffffffu, mine.cnt, src);
      R.qid = __shfl_sync(0xffffffffu, mine.qid, src);
      R.base = shfl_u64(mine.base, src);
      R.qh = R.strand = 0;
      const uint32_t cb = __shfl_sync(0xffffffffu, cov_b, src), ce = __shfl_sync(0xffffffffu, cov_e, src);
      uint32_t* out = c.hit_target + R.base;
      const uint32_t lim = c.capacity > R.base ? (uint32_t)min(c.capacity - R.base, (uint64_t)0xffffffffu) : 0u;
      // (measured and not kept: a predicate-free path for whole 128-row blocks, writing the query id next
      // to each hit instead of the fill below, and 4 rows per lane with 128-bit loads)
      uint32_t count = 0;
      if (ce > cb) {
        LongCtx c2 = c;
        c2.high = a.lc_high;
        c2.ids = a.lc_id;
        c2.n_rows = a.lc_entries;
        LongRange R2 = R;
        R2.lb = cb;
        R2.ub = ce;
        for (uint32_t r0 = cb; r0 < ce; r0 += 32 * kTrips) {
          LongRegs g;
          long_preload<true, false>(c2, lane, R2, r0, g);
          long_consume<true, false>(c2, lane, R2, r0, g, out, lim, count);
        }
      }
      for (uint32_t r0 = R.lb; r0 < R.ub; r0 += 32 * kTrips) {  // (with a block: only when the query leaves its bin)
        LongRegs g;
        long_preload<true, false>(c, lane, R, r0, g);
        long_consume<true, false>(c, lane, R, r0, g, out, lim, count);
      }
      BCU_DEV_ASSERT(count == R.cnt);
      if (c.hit_query) {  // the query-id column of a range is one value: coalesced fill
        uint32_t* qcol = c.hit_query + R.base;
        const uint32_t n = min(R.cnt, lim);
        for (uint32_t k = lane; k < n; k += 32) qcol[k] = R.qid;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Scatter with caller-supplied offsets / any-overlap bit: static tiles, no prefix step.
template <int MODE>
__global__ void __launch_bounds__(kJoinThreads, kJoinMinBlocks) direct_kernel(const JoinArgs a) {
  __shared__ GroupTables tb;
  __shared__ StageBuffers st;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  load_group_tables(a, tb);
  const uint32_t n_tiles = (uint32_t)(((uint64_t)a.n_vq + kCtaTile - 1) / kCtaTile);
  const uint32_t shift_of[kQPT] = {slot_shift(a, 0), slot_shift(a, 1u & (a.n_comp - 1u)),
                                   slot_shift(a, 2u & (a.n_comp - 1u)), slot_shift(a, 3u & (a.n_comp - 1u))};
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t w0 = tile * (uint32_t)kCtaTile + (uint32_t)warp * kWarpTile;
    const uint32_t q0 = w0 + (uint32_t)lane * kQPT;
    uint32_t ql[kQPT], qh[kQPT], qg[kQPT], lb[kQPT], len[kQPT], mask[kQPT];
    load_queries(a, q0, ql, qh, qg);
#pragma unroll
    const uint32_t no_strand[kQPT] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < kQPT; ++j)
      query_bounds<false>(a, tb, q0 + j < a.n_vq, j & (a.n_comp - 1u), shift_of[j], ql[j], qh[j], qg[j], 0u, lb[j], len[j],
                          mask[j]);
    scan_short<false>(a, ql, qh, no_strand, lb, len, mask);
    bool big[kQPT];
    uint32_t ub[kQPT], cnt[kQPT];
    bool lane_big = false;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {
      big[j] = len[j] > kScalarMax;
      lane_big |= big[j];
      ub[j] = big[j] ? exact_upper_bound(a, lb[j], len[j], qh[j]) : lb[j] + len[j];
      cnt[j] = __popc(mask[j]);
    }
    const bool warp_big = __any_sync(0xffffffffu, lane_big);
    if (warp_big) {  // hit counts of the long ranges (the caller's offsets only give per-query totals)
      const uint4 c4 = long_ranges<false, false>(long_ctx(a), lane, pack_bits(big),
                                                 make_uint4(lb[0], lb[1], lb[2], lb[3]),
                                                 make_uint4(ub[0], ub[1], ub[2], ub[3]),
                                                 make_uint4(ql[0], ql[1], ql[2], ql[3]), make_uint4(0, 0, 0, 0), 0, 0,
                                                 0, 0, 0, make_uint4(0, 0, 0, 0), 0);
      const uint32_t cl[kQPT] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
      for (int j = 0; j < kQPT; ++j)
        if (big[j]) cnt[j] = cl[j];
    }
    if (MODE == kModeAny) {  // OR over the slots of a query (they sit in one lane)
      uint32_t acc = 0;
#pragma unroll
      for (int j = kQPT - 1; j >= 0; --j) {
        const uint32_t v = q0 + j;
        acc |= cnt[j];
        if ((v & (a.n_comp - 1u)) == 0) {
          if (v < a.n_vq) a.any[v >> a.comp_shift] = acc ? 1 : 0;
          acc = 0;
        }
      }
      continue;
    }
    uint64_t off[kQPT];
    uint64_t run = 0;
#pragma unroll
    for (int j = 0; j < kQPT; ++j) {  // slot s of a query starts after the hits of its slots < s
      const uint32_t v = q0 + j;
      if ((v & (a.n_comp - 1u)) == 0) run = 0;
      off[j] = (v < a.n_vq) ? a.offsets[v >> a.comp_shift] + run : 0ull;
      run += cnt[j];
    }
    emit_short<true>(a, st, warp, lane, mask, lb, off, 0, w0);
    if (warp_big)
      long_ranges<true, false>(long_ctx(a), lane, pack_bits(big), make_uint4(lb[0], lb[1], lb[2], lb[3]),
                               make_uint4(ub[0], ub[1], ub[2], ub[3]), make_uint4(ql[0], ql[1], ql[2], ql[3]),
                               make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]), off[0], off[1], off[2], off[3], w0,
                               make_uint4(0, 0, 0, 0), 0);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int sm_count(int device) {
  static std::atomic<int> cache[64];
  if (device >= 0 && device < 64) {
    int v = cache[device].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return 148;
  }
  if (device >= 0 && device < 64) cache[device].store(sms, std::memory_order_relaxed);
  return sms;
}

static int launch_dependent(void (*kernel)(const JoinArgs), unsigned grid, cudaStream_t stream,
                            const JoinArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kJoinThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BCU_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  BCU_LAUNCHED();
  return BCU_OK;
}

int launch_join(const bcu_index* ix, int mode, uint64_t n_q, const uint32_t* d_qgroup,
                const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                uint64_t pair_capacity, uint32_t* d_hit_query, uint32_t* d_hit_target,
                uint64_t* d_total, uint8_t* d_any, uint32_t query_id_base, cudaStream_t stream,
                const uint64_t* d_offset_base, const bcu_filter* filter, const uint8_t* d_qstrand,
                uint64_t* total_mapped) {
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
  if (!d_offset_base && !(filter && filter->kind != BCU_FILTER_NONE)) {
    // large batches against an index beyond L2 take the binned path (binned_join.cu) when the index has a bin layout
    const int rc = launch_join_binned(ix, mode, n_q, d_qgroup, d_qlow, d_qhigh, d_offsets, pair_capacity, d_hit_query,
                                      d_hit_target, d_total, query_id_base, stream, total_mapped);
    if (rc != BCU_NOT_TAKEN) return rc;
  }
  JoinArgs a;
  a.lowhigh = ix->d_lowhigh;
  a.high = ix->d_high;
  a.ids = ix->d_id;
  a.dir = ix->d_dir;
  a.hs = ix->d_hs;
  a.dirh = ix->d_dirh;
  a.groups = ix->d_groups;
  a.n_rows = ix->n;
  a.n_dir = ix->n_bins;
  a.n_state = 0;
  a.n_groups = ix->n_groups;
  a.shifts = ix->shifts;
  a.max_gval = ix->max_gval;
  a.qgroup = d_qgroup;
  a.qlow = d_qlow;
  a.qhigh = d_qhigh;
  a.n_q = (uint32_t)n_q;
  a.n_comp = ix->n_comp;
  a.comp_shift = ix->n_comp == 4 ? 2u : (ix->n_comp == 2 ? 1u : 0u);
  const uint64_t n_vq = n_q * ix->n_comp;  // virtual queries: one per (query, length-class slot)
  if (n_vq > 0xfffffffeull) { set_error("query batch x length classes exceeds 2^32-2"); return BCU_E_LIMIT; }
  a.n_vq = (uint32_t)n_vq;
  a.chunk = 0;
  a.vec_ok = aligned16(d_qlow) && aligned16(d_qhigh) && (!d_qgroup || aligned16(d_qgroup)) &&
             (!d_offsets || aligned16(d_offsets));
  a.offsets = d_offsets;
  a.capacity = (mode == kModeScatter) ? ~0ull : pair_capacity;
  a.hit_query = d_hit_query;
  a.hit_target = d_hit_target;
  a.total = d_total;
  a.total_mapped = total_mapped;
  a.any = d_any;
  a.qid_base = query_id_base;
  a.st_lb = nullptr;
  a.st_w = nullptr;
  a.st_ub = nullptr;
  a.cta_total = nullptr;
  a.big_list = nullptr;
  a.cta_big = nullptr;
  a.long_split = 1;
  a.base_in = d_offset_base;
  const bool filt = filter && filter->kind != BCU_FILTER_NONE;
  a.filter_kind = filt ? filter->kind : 0u;
  a.filter_diff = filt ? filter->diff : 0u;
  a.filter_use_strand = filt ? filter->use_strand : 0u;
  a.qstrand = filt ? d_qstrand : nullptr;
  a.lc_seg = ix->lc_bins ? ix->d_lc_seg : nullptr;
  a.lc_off = ix->d_lc_off;
  a.lc_row0 = ix->d_lc_row0;
  a.lc_high = ix->d_lc_high;
  a.lc_id = ix->d_lc_id;
  a.lc_shift = ix->lc_shift;
  a.lc_entries = ix->lc_entries;
  a.lc_bins = ix->lc_bins;
  if (filt && !prefix) { set_error("pair filters are only supported by the count/join entry points"); return BCU_E_INVALID; }
  if (filt && filter->kind != BCU_FILTER_SV2NL_DUP && filter->kind != BCU_FILTER_SV2NL_INV) {
    set_error("unknown pair filter kind %u", filter->kind);
    return BCU_E_INVALID;
  }
  const uint64_t n_tiles = (n_vq + kCtaTile - 1) / kCtaTile;
  // One wave: every CTA of the grid is resident from the start (4 per SM for K3/K4). Measured on B with
  // 1/2/3/4/6 waves: 191/203/198/207/205 us -- a second wave starts staggered behind the first one's
  // stragglers and the per-CTA prologues repeat. BCU_WAVES overrides (finer chunks balance skewed batches).
  static const uint64_t waves = [] {
    const char* e = getenv("BCU_WAVES");
    const long v = e ? atol(e) : 1;
    return (uint64_t)(v < 1 ? 1 : v);
  }();
  const uint64_t cta_budget = (uint64_t)sm_count(ix->device) * kJoinMinBlocks * waves;
  if (!prefix) {
    const unsigned grid = (unsigned)(n_tiles < cta_budget ? n_tiles : cta_budget);
    if (mode == kModeScatter) direct_kernel<kModeScatter><<<grid, kJoinThreads, 0, stream>>>(a);
    else direct_kernel<kModeAny><<<grid, kJoinThreads, 0, stream>>>(a);
    BCU_LAUNCHED();
    return BCU_OK;
  }
  // chunking shared by probe and emit: contiguous runs of whole CTA steps
  const uint64_t tiles_per_cta = (n_tiles + cta_budget - 1) / cta_budget;
  const unsigned grid = (unsigned)((n_tiles + tiles_per_cta - 1) / tiles_per_cta);
  a.chunk = (uint32_t)(tiles_per_cta * kCtaTile);
  const uint64_t padded = n_tiles * kCtaTile;
  void* scratch = nullptr;
  const uint64_t grid_even = ((uint64_t)grid + 1) & ~1ull;  // keeps the state arrays 16-byte aligned
  BCU_CUDA(cudaMallocAsync(&scratch, padded * 16 + grid_even * 12, stream));
  struct ScratchGuard {  // stream-ordered free on every exit path
    void* p;
    cudaStream_t s;
    ~ScratchGuard() { cudaFreeAsync(p, s); }
  } scratch_guard{scratch, stream};
  a.cta_total = reinterpret_cast<uint64_t*>(scratch);
  a.st_lb = reinterpret_cast<uint32_t*>(a.cta_total + grid_even);
  a.st_w = a.st_lb + padded;
  a.st_ub = a.st_w + padded;
  a.big_list = a.st_ub + padded;
  a.cta_big = a.big_list + padded;
  a.n_state = padded;
  if (filt) probe_kernel<true><<<grid, kJoinThreads, 0, stream>>>(a);
  else probe_kernel<false><<<grid, kJoinThreads, 0, stream>>>(a);
  BCU_LAUNCHED();
  // K4 / K4b are launched as programmatic dependents: their CTAs are placed while the previous kernel
  // drains and block in grid_dependency_wait() until it has completed (hides the launch gaps)
  if (mode == kModeFused) {
    if (filt) BCU_TRY(launch_dependent(emit_kernel<true, true>, grid, stream, a));
    else BCU_TRY(launch_dependent(emit_kernel<true, false>, grid, stream, a));
    if (!filt) {  // about two waves of K4b's own occupancy
      static const uint64_t long_waves = [] { const char* e = getenv("BCU_LONG_WAVES"); long v = e ? atol(e) : 2; return (uint64_t)(v < 1 ? 1 : v); }();
      a.long_split = (uint32_t)std::min<uint64_t>(
          16, std::max<uint64_t>(1, ((uint64_t)sm_count(ix->device) * kLongMinBlocks * long_waves + grid - 1) / grid));
      BCU_TRY(launch_dependent(emit_long_kernel, grid * a.long_split, stream, a));
    }
  } else {
    BCU_TRY(launch_dependent(emit_kernel<false, false>, grid, stream, a));
  }
  return BCU_OK;
}

}  // namespace bcu
