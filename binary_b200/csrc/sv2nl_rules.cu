// bcu_sv2nl_join: one sv2nl mapper on the device from the queries to the CSR that is written out.
//
// Reference being replaced (standalone/sv2nl in the reference checkout), per NL record of one mapper:
//   interval_tree.find_overlaps(validated record)                         include/mapper.hpp:214
//   | filter(check_condition)    Dup / Inv: source/mapper.cpp:50-79  (fused into the join kernels: join.cu accept<>)
//                                Tra:       source/mapper.cpp:144-156 (tra_keep below)
//   SV2NL_USE_CACHE: a record whose format_map_key (include/helper.hpp:84-91) was already stored by an EARLIER
//   record with at least one kept hit is not written again          include/mapper.hpp:204-229
// Here: the (filtered) join runs on the device, then the rules kernels of this file work on its CSR in place:
//   1. rules_count_kernel    per record: pairs of its `probes_per_record` consecutive queries that pass tra_keep
//   2. duplicate-key rule    records with hits -> (128-bit key, record) sorted by key with two stable 64-bit LSD
//                            passes (radix_sort.cu; stable = ascending record inside a key) -> every record but the
//                            first of a key run loses its pairs
//   3. exclusive scan of the counts (scan.cu) -> the record-level offsets
//   4. rules_scatter_kernel  kept pairs to their final place
// Only the final offsets and targets travel back to the host.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "common.cuh"

namespace bcu {
namespace {

constexpr int kRulesThreads = 256;

struct RulesArgs {
  uint32_t n_rec, probes;
  uint32_t tra, diff;
  const uint64_t* q_off;    // [n_rec * probes + 1] offsets of the join
  const uint32_t* pairs;    // target ids of the join
  const uint32_t *rec_p1, *rec_p2;                     // tra, per record
  const uint32_t *tgt_p1, *tgt_p2, *tgt_pos, *tgt_end; // tra, per target id
  uint32_t n_t;
};

__device__ __forceinline__ uint32_t absdiff_u32(uint32_t a, uint32_t b) { return a >= b ? a - b : b - a; }

// TraMapper::check_condition (mapper.cpp:144-156) on a pair the re-keyed join produced (equal ordered chromosome
// pairs are the join's group key), and the raw-interval overlap of the reference's find_overlaps on the tree of
// UNVALIDATED BND records (mapper.cpp:103,158-170): the validated NL record has pos = min, svend = max of its two
// breakpoint positions (helper.hpp:52-63).
__device__ __forceinline__ bool tra_keep(const RulesArgs& a, uint32_t rec, uint32_t t) {
  BCU_DEV_ASSERT(t < a.n_t);
  const uint32_t q1 = a.rec_p1[rec], q2 = a.rec_p2[rec];
  if (absdiff_u32(q1, a.tgt_p1[t]) > a.diff || absdiff_u32(q2, a.tgt_p2[t]) > a.diff) return false;
  const uint32_t n_pos = min(q1, q2), n_end = max(q1, q2);
  return n_pos <= a.tgt_end[t] && a.tgt_pos[t] <= n_end;
}

__global__ void __launch_bounds__(kRulesThreads) rules_count_kernel(const RulesArgs a, uint32_t* __restrict__ cnt) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > a.n_rec) return;
  uint32_t c = 0;
  if (r < a.n_rec) {
    const uint64_t b = a.q_off[(uint64_t)r * a.probes], e = a.q_off[(uint64_t)(r + 1) * a.probes];
    if (a.tra) {
      for (uint64_t k = b; k < e; ++k) c += tra_keep(a, r, a.pairs[k]);
    } else {
      c = (uint32_t)(e - b);
    }
  }
  cnt[r] = c;  // cnt[n_rec] = 0: the scan's total slot
}

__global__ void __launch_bounds__(kRulesThreads)
    rules_flag_kernel(const uint32_t* __restrict__ cnt, uint32_t n_rec, uint32_t* __restrict__ flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_rec) flag[r] = (r < n_rec && cnt[r] != 0) ? 1u : 0u;
}

// records with hits, in record order: vals = record, keys = one 64-bit half of its map key
__global__ void __launch_bounds__(kRulesThreads)
    rules_list_kernel(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ pos, uint32_t n_rec,
                      const uint32_t* __restrict__ rec_key, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec || cnt[r] == 0) return;
  const uint32_t i = pos[r];
  vals[i] = r;
  keys[i] = ((uint64_t)rec_key[4 * (uint64_t)r + 2] << 32) | rec_key[4 * (uint64_t)r + 3];  // low half first (LSD)
}
__global__ void __launch_bounds__(kRulesThreads)
    rules_high_key_kernel(const uint32_t* __restrict__ vals, uint32_t m, const uint32_t* __restrict__ rec_key,
                          uint64_t* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t r = vals[i];
  keys[i] = ((uint64_t)rec_key[4 * r] << 32) | rec_key[4 * r + 1];
}
// sorted by key, ascending record inside a key: every record after the first of its run is a repeat
__global__ void __launch_bounds__(kRulesThreads)
    rules_repeat_kernel(const uint32_t* __restrict__ vals, uint32_t m, const uint32_t* __restrict__ rec_key,
                        uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 || i >= m) return;
  const uint4 a = reinterpret_cast<const uint4*>(rec_key)[vals[i]], b = reinterpret_cast<const uint4*>(rec_key)[vals[i - 1]];
  if (a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w) cnt[vals[i]] = 0;
}

__global__ void __launch_bounds__(kRulesThreads)
    rules_scatter_kernel(const RulesArgs a, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off32,
                         uint64_t* __restrict__ off_out, uint32_t* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > a.n_rec) return;
  off_out[r] = off32[r];
  if (r == a.n_rec || cnt[r] == 0) return;
  uint32_t w = off32[r];
  const uint64_t b = a.q_off[(uint64_t)r * a.probes], e = a.q_off[(uint64_t)(r + 1) * a.probes];
  for (uint64_t k = b; k < e; ++k) {
    const uint32_t t = a.pairs[k];
    if (!a.tra || tra_keep(a, r, t)) out[w++] = t;
  }
  BCU_DEV_ASSERT(w == off32[r] + cnt[r]);
}

struct DeviceScratch {  // stream-ordered allocations, freed on every exit path
  cudaStream_t stream = nullptr;
  std::vector<void*> ptrs;
  ~DeviceScratch() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
  }
  int init() {
    BCU_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return BCU_OK;
  }
  template <class T> int alloc(T** p, uint64_t count) {
    void* q = nullptr;
    BCU_CUDA(cudaMallocAsync(&q, std::max<uint64_t>(count, 1) * sizeof(T), stream));
    ptrs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return BCU_OK;
  }
  template <class T> int upload(T** p, const T* host, uint64_t count) {
    *p = nullptr;
    if (!host) return BCU_OK;
    BCU_TRY(alloc(p, count));
    if (count) BCU_CUDA(cudaMemcpyAsync(*p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    return BCU_OK;
  }
};

}  // namespace
}  // namespace bcu

using namespace bcu;

extern "C" int bcu_sv2nl_join(const bcu_index* ix, const bcu_filter* filter, const bcu_sv2nl_rules* rules,
                              uint64_t n_rec, const uint32_t* qgroup, const uint32_t* qlow, const uint32_t* qhigh,
                              const uint8_t* qstrand, uint64_t* offsets, uint64_t pair_capacity,
                              uint32_t* hit_target, uint64_t* total) {
  if (!ix || !rules || !offsets || !total) { set_error("bcu_sv2nl_join: NULL argument"); return BCU_E_INVALID; }
  const uint64_t probes = rules->probes_per_record;
  if (probes == 0 || probes > 16) { set_error("bcu_sv2nl_join: probes_per_record must be 1..16"); return BCU_E_INVALID; }
  const uint64_t n_q = n_rec * probes;
  if (n_q > 0xfffffffeull) { set_error("bcu_sv2nl_join: more than 2^32-2 queries"); return BCU_E_LIMIT; }
  if (n_q && (!qlow || !qhigh)) { set_error("bcu_sv2nl_join: qlow/qhigh are NULL"); return BCU_E_INVALID; }
  if (pair_capacity && !hit_target) { set_error("bcu_sv2nl_join: hit_target is NULL"); return BCU_E_INVALID; }
  if (rules->tra && n_rec && ix->n &&
      (!rules->rec_p1 || !rules->rec_p2 || !rules->tgt_p1 || !rules->tgt_p2 || !rules->tgt_pos || !rules->tgt_end)) {
    set_error("bcu_sv2nl_join: the TRA rule needs the breakpoint columns of records and targets");
    return BCU_E_INVALID;
  }
  if (rules->dedup && n_rec && !rules->rec_key) { set_error("bcu_sv2nl_join: dedup needs rec_key"); return BCU_E_INVALID; }
  const bool filt = filter && filter->kind != BCU_FILTER_NONE;
  if (filt && filter->kind == BCU_FILTER_SV2NL_INV && filter->use_strand && n_q && !qstrand) {
    set_error("bcu_sv2nl_join: the INV filter with use_strand needs qstrand");
    return BCU_E_INVALID;
  }
  *total = 0;
  std::fill(offsets, offsets + n_rec + 1, 0ull);
  if (n_rec == 0 || ix->n == 0) return BCU_OK;
  DeviceGuard guard(ix->device);
  if (!guard.ok) { set_error("bcu_sv2nl_join: cannot select CUDA device %d", ix->device); return BCU_E_CUDA; }
  DeviceScratch s;
  BCU_TRY(s.init());
  cudaStream_t stream = s.stream;
  uint32_t *d_g, *d_l, *d_h;
  uint8_t* d_s;
  uint64_t *d_qoff, *d_total;
  BCU_TRY(s.upload(&d_g, qgroup, n_q));
  BCU_TRY(s.upload(&d_l, qlow, n_q));
  BCU_TRY(s.upload(&d_h, qhigh, n_q));
  BCU_TRY(s.upload(&d_s, qstrand, n_q));
  BCU_TRY(s.alloc(&d_qoff, n_q + 1));
  BCU_TRY(s.alloc(&d_total, 1));
  // ---- the join: count, then count + scatter into an exactly sized buffer -----------------------------------
  BCU_TRY(launch_join(ix, kModeCount, n_q, d_g, d_l, d_h, d_qoff, 0, nullptr, nullptr, d_total, nullptr, 0, stream,
                      nullptr, filt ? filter : nullptr, d_s));
  uint64_t n_pairs = 0;
  BCU_CUDA(cudaMemcpyAsync(&n_pairs, d_qoff + n_q, 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  if (n_pairs >= 0xffffffffull) { set_error("bcu_sv2nl_join: the join has 2^32-1 pairs or more"); return BCU_E_LIMIT; }
  if (n_pairs == 0) return BCU_OK;
  uint32_t* d_pairs;
  BCU_TRY(s.alloc(&d_pairs, n_pairs));
  BCU_TRY(launch_join(ix, kModeFused, n_q, d_g, d_l, d_h, d_qoff, n_pairs, nullptr, d_pairs, d_total, nullptr, 0,
                      stream, nullptr, filt ? filter : nullptr, d_s));
  // ---- the rules ---------------------------------------------------------------------------------------------
  RulesArgs a;
  a.n_rec = (uint32_t)n_rec;
  a.probes = (uint32_t)probes;
  a.tra = rules->tra ? 1u : 0u;
  a.diff = rules->diff;
  a.q_off = d_qoff;
  a.pairs = d_pairs;
  a.n_t = (uint32_t)ix->n;
  uint32_t *d_rp1 = nullptr, *d_rp2 = nullptr, *d_tp1 = nullptr, *d_tp2 = nullptr, *d_tpos = nullptr, *d_tend = nullptr;
  if (a.tra) {
    BCU_TRY(s.upload(&d_rp1, rules->rec_p1, n_rec));
    BCU_TRY(s.upload(&d_rp2, rules->rec_p2, n_rec));
    BCU_TRY(s.upload(&d_tp1, rules->tgt_p1, ix->n));
    BCU_TRY(s.upload(&d_tp2, rules->tgt_p2, ix->n));
    BCU_TRY(s.upload(&d_tpos, rules->tgt_pos, ix->n));
    BCU_TRY(s.upload(&d_tend, rules->tgt_end, ix->n));
  }
  a.rec_p1 = d_rp1; a.rec_p2 = d_rp2; a.tgt_p1 = d_tp1; a.tgt_p2 = d_tp2; a.tgt_pos = d_tpos; a.tgt_end = d_tend;
  uint32_t *d_cnt, *d_off32;
  BCU_TRY(s.alloc(&d_cnt, n_rec + 1));
  BCU_TRY(s.alloc(&d_off32, n_rec + 1));
  const unsigned grid = (unsigned)((n_rec + 1 + kRulesThreads - 1) / kRulesThreads);
  rules_count_kernel<<<grid, kRulesThreads, 0, stream>>>(a, d_cnt);
  BCU_LAUNCHED();
  if (rules->dedup) {
    uint32_t* d_key;
    BCU_TRY(s.upload(&d_key, rules->rec_key, 4 * n_rec));
    rules_flag_kernel<<<grid, kRulesThreads, 0, stream>>>(d_cnt, a.n_rec, d_off32);
    BCU_LAUNCHED();
    uint32_t* d_pos;
    BCU_TRY(s.alloc(&d_pos, n_rec + 1));
    BCU_TRY(exclusive_sum_u32(d_off32, d_pos, n_rec + 1, stream));
    uint32_t m = 0;
    BCU_CUDA(cudaMemcpyAsync(&m, d_pos + n_rec, 4, cudaMemcpyDeviceToHost, stream));
    BCU_CUDA(cudaStreamSynchronize(stream));
    if (m > 1) {
      uint64_t *k_a, *k_b, *k_out;
      uint32_t *v_a, *v_b, *v_out, passes = 0;
      BCU_TRY(s.alloc(&k_a, m));
      BCU_TRY(s.alloc(&k_b, m));
      BCU_TRY(s.alloc(&v_a, m));
      BCU_TRY(s.alloc(&v_b, m));
      rules_list_kernel<<<grid, kRulesThreads, 0, stream>>>(d_cnt, d_pos, a.n_rec, d_key, k_a, v_a);
      BCU_LAUNCHED();
      BCU_TRY(radix_sort_pairs(k_a, k_b, v_a, v_b, m, ~0ull, stream, &k_out, &v_out, &passes));
      uint64_t* k_free = k_out == k_a ? k_b : k_a;  // the high halves go where the sorted low halves are not
      uint32_t* v_free = v_out == v_a ? v_b : v_a;
      const unsigned mgrid = (unsigned)((m + kRulesThreads - 1) / kRulesThreads);
      rules_high_key_kernel<<<mgrid, kRulesThreads, 0, stream>>>(v_out, m, d_key, k_free);
      BCU_LAUNCHED();
      uint64_t* k2;
      uint32_t* v2;
      BCU_TRY(radix_sort_pairs(k_free, k_out, v_out, v_free, m, ~0ull, stream, &k2, &v2, &passes));
      rules_repeat_kernel<<<mgrid, kRulesThreads, 0, stream>>>(v2, m, d_key, d_cnt);
      BCU_LAUNCHED();
    }
  }
  BCU_TRY(exclusive_sum_u32(d_cnt, d_off32, n_rec + 1, stream));
  uint32_t* d_out;
  uint64_t* d_off_out;
  BCU_TRY(s.alloc(&d_out, n_pairs));
  BCU_TRY(s.alloc(&d_off_out, n_rec + 1));
  rules_scatter_kernel<<<grid, kRulesThreads, 0, stream>>>(a, d_cnt, d_off32, d_off_out, d_out);
  BCU_LAUNCHED();
  BCU_CUDA(cudaMemcpyAsync(offsets, d_off_out, (n_rec + 1) * 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  *total = offsets[n_rec];
  if (*total > pair_capacity) {
    set_error("bcu_sv2nl_join: %llu pairs, capacity %llu", (unsigned long long)*total, (unsigned long long)pair_capacity);
    return BCU_E_CAPACITY;
  }
  if (*total) {
    BCU_CUDA(cudaMemcpyAsync(hit_target, d_out, *total * 4, cudaMemcpyDeviceToHost, stream));
    BCU_CUDA(cudaStreamSynchronize(stream));
  }
  return BCU_OK;
}
