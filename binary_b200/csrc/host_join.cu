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
#include <string>
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

__global__ void counts_from_offsets_kernel(const uint64_t* __restrict__ off, uint64_t n, uint32_t* __restrict__ counts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = (uint32_t)(off[i + 1] - off[i]);
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

// offsets (u64, n_q + 1) and/or counts (u32 hits per query, n_q: half the bytes over PCIe; plain joins only)
static int join_host(const bcu_index* ix, const bcu_filter* filter, uint64_t n_q, const uint32_t* qgroup,
                     const uint32_t* qlow, const uint32_t* qhigh, const uint8_t* qstrand, uint64_t* offsets,
                     uint64_t pair_capacity, uint32_t* hit_query, uint32_t* hit_target, uint64_t* total,
                     uint32_t* counts = nullptr) {
  if (!ix) { set_error("bcu_join: index is NULL"); return BCU_E_INVALID; }
  if (n_q && (!qlow || !qhigh)) { set_error("bcu_join: qlow/qhigh are NULL"); return BCU_E_INVALID; }
  if (n_q > 0xfffffffeull) { set_error("bcu_join: n_q exceeds 2^32-2"); return BCU_E_LIMIT; }
  if ((!offsets && !counts) || !total) { set_error("bcu_join: offsets/total are NULL"); return BCU_E_INVALID; }
  if (counts && qstrand) { set_error("bcu_join: per-query counts are not available with a strand filter"); return BCU_E_INVALID; }
  if (pair_capacity && !hit_target) {
    set_error("bcu_join: hit_target is NULL");
    return BCU_E_INVALID;
  }
  *total = 0;
  if (n_q == 0 || ix->n == 0) {
    if (offsets) std::fill(offsets, offsets + n_q + 1, 0ull);
    if (counts) std::fill(counts, counts + n_q, 0u);
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
  if (counts) BCU_TRY(c->grow(&c->d_qs, &c->cap_qs, n_q * 4));  // the strand column's buffer doubles as the u32 counts

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
    if (counts) {  // the chunk's last offset is the next chunk's first: take it from the running total
      BCU_CUDA(cudaMemcpyAsync(c->d_off + b + n, c->d_totals + i, 8, cudaMemcpyDeviceToDevice, c->s_run));
      counts_from_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->s_run>>>(
          c->d_off + b, n, reinterpret_cast<uint32_t*>(c->d_qs) + b);
      BCU_LAUNCHED();
    }
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
    if (offsets) BCU_CUDA(cudaMemcpyAsync(offsets + b, c->d_off + b, n_off * 8, cudaMemcpyDeviceToHost, c->s_out));
    if (counts)
      BCU_CUDA(cudaMemcpyAsync(counts + b, reinterpret_cast<uint32_t*>(c->d_qs) + b, n * 4, cudaMemcpyDeviceToHost, c->s_out));
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

// ---------------------------------------------------------------------------------------------------
// bcu_join_multi: ONE call, several GPUs (the reference's counterpart is one pool task per
// chromosome, sv2nl mapper.hpp:238-246, over one shared tree, mapper.cpp:136-140). The batch is cut into
// n_dev contiguous query ranges; device d joins range d against indexes[d] (replicas of the same target
// set). No device ever talks to another: the only exchange is the per-range hit total, on the host.
//   phase 1  every device: queries of its range H2D, count pass -> the range's total
//   (host)   exclusive sum of the totals -> where every range's pairs start in the caller's buffers
//   phase 2  every device: the chunked join of host_join.cu -- kernels overlapped with the D2H copies of
//            offsets and pairs -- with the range's base added on the device, written straight to their final
//            place in the caller's buffers
// One persistent worker thread per device (created on first use) owns that device's streams and staging.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace bcu {

class DeviceWorker {  // a thread that runs one job at a time on one device
 public:
  explicit DeviceWorker(int device) : device_(device), thread_([this] { loop(); }) {}
  ~DeviceWorker() {
    { std::lock_guard<std::mutex> l(m_); quit_ = true; }
    cv_.notify_all();
    thread_.join();
  }
  void submit(std::function<int()> job) {
    { std::lock_guard<std::mutex> l(m_); job_ = std::move(job); has_job_ = true; done_ = false; }
    cv_.notify_all();
  }
  int wait(std::string* err) {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [this] { return done_; });
    if (err) *err = error_;
    return rc_;
  }

 private:
  void loop() {
    cudaSetDevice(device_);
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return has_job_ || quit_; });
        if (quit_) return;
        job = std::move(job_);
        has_job_ = false;
      }
      const int rc = job();
      {
        std::lock_guard<std::mutex> l(m_);
        rc_ = rc;
        error_ = rc == BCU_OK ? "" : bcu_last_error();  // the message lives in this thread's slot: hand it over
        done_ = true;
      }
      cv_.notify_all();
    }
  }
  int device_;
  std::mutex m_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool has_job_ = false, done_ = true, quit_ = false;
  int rc_ = BCU_OK;
  std::string error_;
  std::thread thread_;
};

static DeviceWorker* worker_for(int device) {
  static std::mutex m;
  static std::unique_ptr<DeviceWorker> workers[kMaxDevices];
  if (device < 0 || device >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> l(m);
  if (!workers[device]) workers[device].reset(new (std::nothrow) DeviceWorker(device));
  return workers[device].get();
}

struct MultiRange {
  const bcu_index* ix;
  uint64_t begin, n;       // query range
  uint64_t total = 0;      // hits of the range (phase 1)
  uint64_t base = 0;       // hits of all earlier ranges (host)
};

// phase 1 on the worker of r.ix->device
static int multi_count(MultiRange& r, const uint32_t* qgroup, const uint32_t* qlow, const uint32_t* qhigh) {
  if (r.n == 0 || r.ix->n == 0) { r.total = 0; return BCU_OK; }
  HostCtx* c = host_ctx(r.ix->device);
  if (!c) { set_error("bcu_join_multi: host context allocation failed"); return BCU_E_NOMEM; }
  BCU_TRY(c->prepare(r.ix->device, r.n, 0, 1, qgroup != nullptr, false, false));
  if (qgroup) BCU_CUDA(cudaMemcpyAsync(c->d_qg, qgroup + r.begin, r.n * 4, cudaMemcpyHostToDevice, c->s_run));
  BCU_CUDA(cudaMemcpyAsync(c->d_ql, qlow + r.begin, r.n * 4, cudaMemcpyHostToDevice, c->s_run));
  BCU_CUDA(cudaMemcpyAsync(c->d_qh, qhigh + r.begin, r.n * 4, cudaMemcpyHostToDevice, c->s_run));
  BCU_TRY(launch_join(r.ix, kModeCount, r.n, qgroup ? c->d_qg : nullptr, c->d_ql, c->d_qh, c->d_off, 0, nullptr, nullptr,
                      nullptr, nullptr, 0, c->s_run, nullptr, nullptr, nullptr, c->h_totals_dev));
  BCU_CUDA(cudaStreamSynchronize(c->s_run));
  r.total = c->h_totals[0];
  return BCU_OK;
}

// phase 2: the queries already are on the device; join chunk by chunk, copy out while the next chunk runs
static int multi_join(const MultiRange& r, bool has_group, uint64_t* offsets, uint32_t* counts, uint64_t pair_capacity,
                      uint32_t* hit_query, uint32_t* hit_target) {
  if (r.n == 0) return BCU_OK;
  if (r.ix->n == 0 || r.total == 0) {  // nothing hits: the offsets of the range all equal its base
    if (offsets) std::fill(offsets + r.begin, offsets + r.begin + r.n, r.base);
    if (counts) std::fill(counts + r.begin, counts + r.begin + r.n, 0u);
    return BCU_OK;
  }
  HostCtx* c = host_ctx(r.ix->device);
  const uint64_t local_cap = pair_capacity > r.base ? std::min(pair_capacity - r.base, r.total) : 0;
  const uint64_t kChunk = host_chunk();
  const uint64_t n_chunks = (r.n + kChunk - 1) / kChunk;
  BCU_TRY(c->prepare(r.ix->device, r.n, std::max<uint64_t>(local_cap, 1), n_chunks + 1, has_group, hit_query != nullptr, false));
  if (counts) BCU_TRY(c->grow(&c->d_qs, &c->cap_qs, r.n * 4));  // the strand column's buffer doubles as the u32 counts
  // chunk i starts from the running total of chunk i - 1; slot 0 holds the range's base
  BCU_CUDA(cudaMemcpyAsync(c->d_totals, &r.base, 8, cudaMemcpyHostToDevice, c->s_run));
  // 1. queue every chunk's kernels; nothing here blocks the host
  for (uint64_t i = 0; i < n_chunks; ++i) {
    const uint64_t b = i * kChunk, n = std::min(kChunk, r.n - b);
    // positions are absolute (they include the range's base) while the device buffers hold this range's pairs
    // only: the buffer pointers are moved back by the base
    BCU_TRY(launch_join(r.ix, kModeFused, n, has_group ? c->d_qg + b : nullptr, c->d_ql + b, c->d_qh + b, c->d_off + b,
                        r.base + local_cap, hit_query ? c->d_hq - r.base : nullptr, c->d_ht - r.base, c->d_totals + i + 1,
                        nullptr, (uint32_t)(r.begin + b), c->s_run, c->d_totals + i, nullptr, nullptr, c->h_totals_dev + i + 1));
    if (counts) {  // the chunk's last offset is the next chunk's first: take it from the running total
      BCU_CUDA(cudaMemcpyAsync(c->d_off + b + n, c->d_totals + i + 1, 8, cudaMemcpyDeviceToDevice, c->s_run));
      counts_from_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->s_run>>>(
          c->d_off + b, n, reinterpret_cast<uint32_t*>(c->d_qs) + b);
      BCU_LAUNCHED();
    }
    BCU_CUDA(cudaEventRecord(c->ev_run[i], c->s_run));
  }
  // 2. as each chunk finishes, its running total tells how many pairs to bring back
  uint64_t done_pairs = 0;
  for (uint64_t i = 0; i < n_chunks; ++i) {
    const uint64_t b = i * kChunk, n = std::min(kChunk, r.n - b);
    BCU_CUDA(cudaEventSynchronize(c->ev_run[i]));
    BCU_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_run[i], 0));
    if (offsets) BCU_CUDA(cudaMemcpyAsync(offsets + r.begin + b, c->d_off + b, n * 8, cudaMemcpyDeviceToHost, c->s_out));
    if (counts)
      BCU_CUDA(cudaMemcpyAsync(counts + r.begin + b, reinterpret_cast<uint32_t*>(c->d_qs) + b, n * 4, cudaMemcpyDeviceToHost, c->s_out));
    const uint64_t upto = std::min(c->h_totals[i + 1] - r.base, local_cap);
    if (upto > done_pairs) {
      if (hit_query)
        BCU_CUDA(cudaMemcpyAsync(hit_query + r.base + done_pairs, c->d_hq + done_pairs, (upto - done_pairs) * 4,
                                 cudaMemcpyDeviceToHost, c->s_out));
      BCU_CUDA(cudaMemcpyAsync(hit_target + r.base + done_pairs, c->d_ht + done_pairs, (upto - done_pairs) * 4,
                               cudaMemcpyDeviceToHost, c->s_out));
      done_pairs = upto;
    }
  }
  BCU_CUDA(cudaStreamSynchronize(c->s_out));
  return BCU_OK;
}

}  // namespace bcu

extern "C" int bcu_join_multi(const bcu_index* const* indexes, int n_dev, uint64_t n_q, const uint32_t* qgroup,
                              const uint32_t* qlow, const uint32_t* qhigh, uint64_t* offsets, uint32_t* counts,
                              uint64_t pair_capacity, uint32_t* hit_query, uint32_t* hit_target, uint64_t* total) {
  if (!indexes || n_dev < 1 || n_dev > kMaxDevices) { set_error("bcu_join_multi: bad index list"); return BCU_E_INVALID; }
  if (n_q && (!qlow || !qhigh)) { set_error("bcu_join_multi: qlow/qhigh are NULL"); return BCU_E_INVALID; }
  if (n_q > 0xfffffffeull) { set_error("bcu_join_multi: n_q exceeds 2^32-2"); return BCU_E_LIMIT; }
  if ((!offsets && !counts) || !total) { set_error("bcu_join_multi: offsets/counts and total are NULL"); return BCU_E_INVALID; }
  if (pair_capacity && !hit_target) { set_error("bcu_join_multi: hit_target is NULL"); return BCU_E_INVALID; }
  for (int d = 0; d < n_dev; ++d) {
    if (!indexes[d]) { set_error("bcu_join_multi: index %d is NULL", d); return BCU_E_INVALID; }
    if (indexes[d]->n != indexes[0]->n) { set_error("bcu_join_multi: the indexes are not replicas of one target set"); return BCU_E_INVALID; }
    for (int e = 0; e < d; ++e)
      if (indexes[e]->device == indexes[d]->device) { set_error("bcu_join_multi: two indexes on device %d", indexes[d]->device); return BCU_E_INVALID; }
  }
  *total = 0;
  if (n_dev == 1)  // no range needs another's total: the full-duplex chunk pipeline of bcu_join, with the counts option
    return join_host(indexes[0], nullptr, n_q, qgroup, qlow, qhigh, nullptr, offsets, pair_capacity, hit_query, hit_target,
                     total, counts);
  std::vector<MultiRange> ranges(n_dev);
  std::vector<DeviceWorker*> workers(n_dev);
  const uint64_t base_n = n_q / n_dev, extra = n_q % n_dev;  // binary_b200.sharding.shard_range
  for (int d = 0; d < n_dev; ++d) {
    ranges[d].ix = indexes[d];
    ranges[d].begin = d * base_n + std::min<uint64_t>(d, extra);
    ranges[d].n = base_n + ((uint64_t)d < extra ? 1 : 0);
    workers[d] = worker_for(indexes[d]->device);
    if (!workers[d]) { set_error("bcu_join_multi: cannot start the worker of device %d", indexes[d]->device); return BCU_E_NOMEM; }
  }
  auto run_phase = [&](const std::function<int(int)>& job) -> int {
    for (int d = 0; d < n_dev; ++d) workers[d]->submit([&job, d] { return job(d); });
    int rc = BCU_OK;
    for (int d = 0; d < n_dev; ++d) {
      std::string err;
      const int r = workers[d]->wait(&err);
      if (r != BCU_OK && rc == BCU_OK) { rc = r; set_error("%s", err.c_str()); }
    }
    return rc;
  };
  BCU_TRY(run_phase([&](int d) { return multi_count(ranges[d], qgroup, qlow, qhigh); }));
  uint64_t sum = 0;
  for (int d = 0; d < n_dev; ++d) { ranges[d].base = sum; sum += ranges[d].total; }
  *total = sum;
  BCU_TRY(run_phase([&](int d) {
    return multi_join(ranges[d], qgroup != nullptr, offsets, counts, pair_capacity, hit_query, hit_target);
  }));
  if (offsets) offsets[n_q] = sum;
  if (sum > pair_capacity) {
    set_error("bcu_join_multi: %llu pairs exceed pair_capacity %llu", (unsigned long long)sum, (unsigned long long)pair_capacity);
    return BCU_E_CAPACITY;
  }
  return BCU_OK;
}
