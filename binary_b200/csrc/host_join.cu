// bcu_join with HOST buffers: the end-to-end path a caller of the reference's find_overlaps loop
// (sv2nl mapper.hpp:207-218) switches to. The batch is cut into chunks that flow through three streams
//   copy-in (H2D queries)  ->  run (probe + emit kernels, join.cu)  ->  copy-out (D2H offsets + pairs)
// so PCIe transfers in both directions overlap the kernels. Offsets are global across chunks: each
// chunk's emit kernel starts from the previous chunk's running total, which stays on the device.
// Device staging buffers and streams are cached per host thread and device (grow-only; bcu_trim frees).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "common.cuh"

namespace bcu {

constexpr uint64_t kHostChunkDefault = 2u << 20;  // queries per pipeline chunk (multiple of the CTA step)
constexpr int kMaxDevices = 64;

struct HostCtx {
  int device = -1;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  uint32_t *d_qg = nullptr, *d_ql = nullptr, *d_qh = nullptr;
  uint64_t cap_qg = 0, cap_ql = 0, cap_qh = 0;
  uint8_t* d_qs = nullptr;  // per-query strand bits of a filtered join
  uint64_t cap_qs = 0;
  uint64_t* d_off = nullptr;
  uint64_t cap_off = 0;
  uint32_t *d_hq = nullptr, *d_ht = nullptr;
  uint64_t cap_hq = 0, cap_ht = 0;
  uint64_t* d_totals = nullptr;
  uint64_t* h_totals = nullptr;  // pinned and mapped: the emit kernel writes each chunk's total straight into it
  uint64_t* h_totals_dev = nullptr;  // its device-side address
  uint64_t cap_chunks = 0;
  std::vector<cudaEvent_t> ev_in, ev_run;

  void release() {
    cudaFree(d_qg); cudaFree(d_ql); cudaFree(d_qh); cudaFree(d_qs); cudaFree(d_off); cudaFree(d_hq); cudaFree(d_ht);
    cudaFree(d_totals);
    if (h_totals) cudaFreeHost(h_totals);
    d_qg = d_ql = d_qh = d_hq = d_ht = nullptr;
    d_qs = nullptr;
    cap_qs = 0;
    d_off = d_totals = h_totals = nullptr;
    cap_qg = cap_ql = cap_qh = cap_off = cap_hq = cap_ht = cap_chunks = 0;
    for (cudaEvent_t e : ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_run) cudaEventDestroy(e);
    ev_in.clear();
    ev_run.clear();
    if (s_in) cudaStreamDestroy(s_in);
    if (s_run) cudaStreamDestroy(s_run);
    if (s_out) cudaStreamDestroy(s_out);
    s_in = s_run = s_out = nullptr;
    cudaGetLastError();
  }
  ~HostCtx() { release(); }

  template <class T> int grow(T** p, uint64_t* cap, uint64_t need) {
    if (need <= *cap) return BCU_OK;
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const uint64_t want = need + need / 4;  // headroom: batches of similar size do not reallocate
    BCU_CUDA(cudaMalloc((void**)p, std::max<uint64_t>(want, 1) * sizeof(T)));
    *cap = want;
    return BCU_OK;
  }

  int prepare(int dev, uint64_t n_q, uint64_t pair_capacity, uint64_t n_chunks, bool has_group,
              bool want_query_ids, bool has_strand) {
    device = dev;
    if (!s_in) {
      BCU_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
      BCU_CUDA(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
      BCU_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    }
    if (has_group) BCU_TRY(grow(&d_qg, &cap_qg, n_q));
    if (has_strand) BCU_TRY(grow(&d_qs, &cap_qs, n_q));
    BCU_TRY(grow(&d_ql, &cap_ql, n_q));
    BCU_TRY(grow(&d_qh, &cap_qh, n_q));
    BCU_TRY(grow(&d_off, &cap_off, n_q + 1));
    if (want_query_ids) BCU_TRY(grow(&d_hq, &cap_hq, pair_capacity));
    BCU_TRY(grow(&d_ht, &cap_ht, pair_capacity));
    if (n_chunks > cap_chunks) {
      cudaFree(d_totals);
      if (h_totals) cudaFreeHost(h_totals);
      d_totals = h_totals = nullptr;
      cap_chunks = 0;
      BCU_CUDA(cudaMalloc((void**)&d_totals, n_chunks * 8));
      BCU_CUDA(cudaHostAlloc((void**)&h_totals, n_chunks * 8, cudaHostAllocMapped));
      BCU_CUDA(cudaHostGetDevicePointer((void**)&h_totals_dev, h_totals, 0));
      cap_chunks = n_chunks;
    }
    while (ev_in.size() < n_chunks) {
      cudaEvent_t a, b;
      BCU_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      BCU_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      ev_in.push_back(a);
      ev_run.push_back(b);
    }
    return BCU_OK;
  }
};

// BCU_HOST_CHUNK (queries, rounded to a multiple of 1024) overrides the pipeline granularity
static uint64_t host_chunk() {
  static const uint64_t v = [] {
    const char* e = std::getenv("BCU_HOST_CHUNK");
    uint64_t c = e ? std::strtoull(e, nullptr, 10) : 0;
    if (c < 1024) c = kHostChunkDefault;
    return (c + 1023) / 1024 * 1024;
  }();
  return v;
}

static HostCtx* host_ctx(int device) {
  static thread_local std::unique_ptr<HostCtx> ctx[kMaxDevices];
  if (device < 0 || device >= kMaxDevices) return nullptr;
  if (!ctx[device]) ctx[device].reset(new (std::nothrow) HostCtx());
  return ctx[device].get();
}

void trim_host_ctx() {
  for (int d = 0; d < kMaxDevices; ++d) {
    HostCtx* c = host_ctx(d);
    if (c && c->device >= 0) {
      DeviceGuard guard(c->device);
      c->release();
    }
  }
}

}  // namespace bcu

using namespace bcu;

extern "C" int bcu_trim(void) {
  trim_host_ctx();
  return BCU_OK;
}

static int join_host(const bcu_index* ix, const bcu_filter* filter, uint64_t n_q, const uint32_t* qgroup,
                     const uint32_t* qlow, const uint32_t* qhigh, const uint8_t* qstrand, uint64_t* offsets,
                     uint64_t pair_capacity, uint32_t* hit_query, uint32_t* hit_target, uint64_t* total) {
  if (!ix) { set_error("bcu_join: index is NULL"); return BCU_E_INVALID; }
  if (n_q && (!qlow || !qhigh)) { set_error("bcu_join: qlow/qhigh are NULL"); return BCU_E_INVALID; }
  if (n_q > 0xfffffffeull) { set_error("bcu_join: n_q exceeds 2^32-2"); return BCU_E_LIMIT; }
  if (!offsets || !total) { set_error("bcu_join: offsets/total are NULL"); return BCU_E_INVALID; }
  if (pair_capacity && !hit_target) {
    set_error("bcu_join: hit_target is NULL");
    return BCU_E_INVALID;
  }
  *total = 0;
  if (n_q == 0 || ix->n == 0) {
    std::fill(offsets, offsets + n_q + 1, 0ull);
    return BCU_OK;
  }
  DeviceGuard guard(ix->device);
  if (!guard.ok) { set_error("bcu_join: cannot select CUDA device %d", ix->device); return BCU_E_CUDA; }
  HostCtx* c = host_ctx(ix->device);
  if (!c) { set_error("bcu_join: host context allocation failed"); return BCU_E_NOMEM; }
  // chunk boundaries: the first two chunks are a quarter and a half of the steady size so that the
  // copy-out stream starts early (the pipeline's fill and drain are what is not overlapped)
  const uint64_t kHostChunk = host_chunk();
  std::vector<uint64_t> bounds{0};
  for (uint64_t step : {kHostChunk / 4, kHostChunk / 2}) {
    step = std::max<uint64_t>(step / 1024 * 1024, 1024);
    if (bounds.back() + step < n_q) bounds.push_back(bounds.back() + step);
  }
  // ... and the last two are a half and a quarter: what follows the last H2D copy (its join and its
  // D2H) is not overlapped with anything either
  const uint64_t tail_half = std::max<uint64_t>(kHostChunk / 2 / 1024 * 1024, 1024);
  const uint64_t tail_quarter = std::max<uint64_t>(kHostChunk / 4 / 1024 * 1024, 1024);
  const uint64_t tail = tail_half + tail_quarter;
  while (bounds.back() + kHostChunk + tail < n_q) bounds.push_back(bounds.back() + kHostChunk);
  if (bounds.back() + tail < n_q) {
    const uint64_t t0 = (n_q - tail) / 1024 * 1024;  // chunk starts stay multiples of 1024 (128-bit paths)
    if (t0 > bounds.back()) bounds.push_back(t0);
    bounds.push_back(t0 + tail_half);
  }
  bounds.push_back(n_q);
  const uint64_t n_chunks = bounds.size() - 1;
  BCU_TRY(c->prepare(ix->device, n_q, pair_capacity, n_chunks, qgroup != nullptr, hit_query != nullptr,
                     qstrand != nullptr));

  const bool trace = std::getenv("BCU_HOST_TRACE") != nullptr;  // dev aid: host-side timeline on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  // 1. queue every chunk's H2D copies and kernels; nothing here blocks the host
  for (uint64_t i = 0; i < n_chunks; ++i) {
    const uint64_t b = bounds[i], n = bounds[i + 1] - b;
    if (qgroup) BCU_CUDA(cudaMemcpyAsync(c->d_qg + b, qgroup + b, n * 4, cudaMemcpyHostToDevice, c->s_in));
    BCU_CUDA(cudaMemcpyAsync(c->d_ql + b, qlow + b, n * 4, cudaMemcpyHostToDevice, c->s_in));
    BCU_CUDA(cudaMemcpyAsync(c->d_qh + b, qhigh + b, n * 4, cudaMemcpyHostToDevice, c->s_in));
    if (qstrand) BCU_CUDA(cudaMemcpyAsync(c->d_qs + b, qstrand + b, n, cudaMemcpyHostToDevice, c->s_in));
    BCU_CUDA(cudaEventRecord(c->ev_in[i], c->s_in));
    BCU_CUDA(cudaStreamWaitEvent(c->s_run, c->ev_in[i], 0));
    BCU_TRY(launch_join(ix, kModeFused, n, qgroup ? c->d_qg + b : nullptr, c->d_ql + b, c->d_qh + b,
                        c->d_off + b, pair_capacity, hit_query ? c->d_hq : nullptr, c->d_ht, c->d_totals + i, nullptr,
                        (uint32_t)b, c->s_run, i ? c->d_totals + (i - 1) : nullptr, filter,
                        qstrand ? c->d_qs + b : nullptr, c->h_totals_dev + i));
    // (no D2H copy of the total on s_run: it would queue behind s_out's large copies in the copy engine and
    // stall the next chunk's kernels; the kernel writes the total into mapped host memory instead)
    BCU_CUDA(cudaEventRecord(c->ev_run[i], c->s_run));
    if (trace) fprintf(stderr, "[bcu_join] chunk %llu (%llu queries) queued at %.0f us\n", (unsigned long long)i,
                       (unsigned long long)n, since());
  }
  // 2. as each chunk finishes, its running total tells how many pairs to bring back
  uint64_t done_pairs = 0;
  for (uint64_t i = 0; i < n_chunks; ++i) {
    const uint64_t b = bounds[i], n = bounds[i + 1] - b;
    BCU_CUDA(cudaEventSynchronize(c->ev_run[i]));
    const uint64_t t = c->h_totals[i];
    if (trace) fprintf(stderr, "[bcu_join] chunk %llu joined at %.0f us, running total %llu\n", (unsigned long long)i,
                       since(), (unsigned long long)t);
    BCU_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_run[i], 0));
    const uint64_t n_off = n + (i + 1 == n_chunks ? 1 : 0);
    BCU_CUDA(cudaMemcpyAsync(offsets + b, c->d_off + b, n_off * 8, cudaMemcpyDeviceToHost, c->s_out));
    const uint64_t upto = std::min(t, pair_capacity);
    if (upto > done_pairs) {
      if (hit_query)
        BCU_CUDA(cudaMemcpyAsync(hit_query + done_pairs, c->d_hq + done_pairs, (upto - done_pairs) * 4,
                                 cudaMemcpyDeviceToHost, c->s_out));
      BCU_CUDA(cudaMemcpyAsync(hit_target + done_pairs, c->d_ht + done_pairs, (upto - done_pairs) * 4,
                               cudaMemcpyDeviceToHost, c->s_out));
      done_pairs = upto;
    }
    *total = t;
  }
  BCU_CUDA(cudaStreamSynchronize(c->s_out));
  if (trace) fprintf(stderr, "[bcu_join] last byte on the host at %.0f us\n", since());
  if (*total > pair_capacity) {
    set_error("bcu_join: %llu pairs exceed pair_capacity %llu", (unsigned long long)*total,
              (unsigned long long)pair_capacity);
    return BCU_E_CAPACITY;
  }
  return BCU_OK;
}

extern "C" int bcu_join(const bcu_index* ix, uint64_t n_q, const uint32_t* qgroup, const uint32_t* qlow,
                        const uint32_t* qhigh, uint64_t* offsets, uint64_t pair_capacity,
                        uint32_t* hit_query, uint32_t* hit_target, uint64_t* total) {
  return join_host(ix, nullptr, n_q, qgroup, qlow, qhigh, nullptr, offsets, pair_capacity, hit_query, hit_target,
                   total);
}

extern "C" int bcu_join_filtered(const bcu_index* ix, const bcu_filter* filter, uint64_t n_q,
                                 const uint32_t* qgroup, const uint32_t* qlow, const uint32_t* qhigh,
                                 const uint8_t* qstrand, uint64_t* offsets, uint64_t pair_capacity,
                                 uint32_t* hit_query, uint32_t* hit_target, uint64_t* total) {
  if (!filter) { set_error("bcu_join_filtered: filter is NULL"); return BCU_E_INVALID; }
  if (filter->kind == BCU_FILTER_SV2NL_INV && filter->use_strand && n_q && !qstrand) {
    set_error("bcu_join_filtered: the INV filter with use_strand needs qstrand");
    return BCU_E_INVALID;
  }
  return join_host(ix, filter, n_q, qgroup, qlow, qhigh, qstrand, offsets, pair_capacity, hit_query, hit_target,
                   total);
}
