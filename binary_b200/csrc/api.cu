// extern "C" surface of libbinary_cuda (see include/binary_cuda.h for the contract and the reference
// interfaces each entry point replaces). Host-buffer entry points stage through device memory; the
// *_dev entry points launch directly on the caller's stream.
#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"

namespace bcu {

std::atomic<uint64_t> g_launches{0};
static thread_local char tls_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
}

// Per-call staging: a private non-blocking stream plus device buffers that are released on every
// exit path.
struct Staging {
  cudaStream_t stream = nullptr;
  void* ptrs[16];
  int n = 0;
  ~Staging() {
    for (int i = 0; i < n; ++i) cudaFree(ptrs[i]);
    if (stream) cudaStreamDestroy(stream);
  }
  int init() {
    BCU_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return BCU_OK;
  }
  template <class T> int alloc(T** p, uint64_t count) {
    void* q = nullptr;
    BCU_CUDA(cudaMalloc(&q, std::max<uint64_t>(count, 1) * sizeof(T)));
    ptrs[n++] = q;
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

static int guard_ok(const DeviceGuard& g, int device) {
  if (g.ok) return BCU_OK;
  set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
  return BCU_E_CUDA;
}

static int check_query_args(const char* fn, const bcu_index* ix, uint64_t n_q, const uint32_t* qlow,
                            const uint32_t* qhigh) {
  if (!ix) { set_error("%s: index is NULL", fn); return BCU_E_INVALID; }
  if (n_q && (!qlow || !qhigh)) { set_error("%s: qlow/qhigh are NULL", fn); return BCU_E_INVALID; }
  if (n_q > 0xfffffffeull) { set_error("%s: n_q exceeds 2^32-2", fn); return BCU_E_LIMIT; }
  return BCU_OK;
}

}  // namespace bcu

using namespace bcu;

extern "C" const char* bcu_version(void) { return "binary_b200 libbinary_cuda 0.1 (sm_100a)"; }
extern "C" const char* bcu_last_error(void) { return tls_error; }
extern "C" uint64_t bcu_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int bcu_device_count(int* n) {
  if (!n) { set_error("bcu_device_count: NULL argument"); return BCU_E_INVALID; }
  *n = 0;
  BCU_CUDA(cudaGetDeviceCount(n));
  return BCU_OK;
}

extern "C" int bcu_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) { set_error("bcu_host_alloc: NULL argument"); return BCU_E_INVALID; }
  *ptr = nullptr;
  BCU_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  return BCU_OK;
}

extern "C" int bcu_host_free(void* ptr) {
  if (ptr) BCU_CUDA(cudaFreeHost(ptr));
  return BCU_OK;
}

// ---- device-pointer entry points -------------------------------------------------------------------
extern "C" int bcu_query_count_dev(const bcu_index* ix, uint64_t n_q, const uint32_t* d_qgroup,
                                   const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                                   void* stream) {
  BCU_TRY(check_query_args("bcu_query_count_dev", ix, n_q, d_qlow, d_qhigh));
  if (!d_offsets) { set_error("bcu_query_count_dev: d_offsets is NULL"); return BCU_E_INVALID; }
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  return launch_join(ix, kModeCount, n_q, d_qgroup, d_qlow, d_qhigh, d_offsets, 0, nullptr, nullptr,
                     nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int bcu_query_scatter_dev(const bcu_index* ix, uint64_t n_q, const uint32_t* d_qgroup,
                                     const uint32_t* d_qlow, const uint32_t* d_qhigh,
                                     const uint64_t* d_offsets, uint32_t* d_hit_query,
                                     uint32_t* d_hit_target, void* stream) {
  BCU_TRY(check_query_args("bcu_query_scatter_dev", ix, n_q, d_qlow, d_qhigh));
  if (!d_offsets || !d_hit_query || !d_hit_target) {
    set_error("bcu_query_scatter_dev: NULL output/offset pointer");
    return BCU_E_INVALID;
  }
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  return launch_join(ix, kModeScatter, n_q, d_qgroup, d_qlow, d_qhigh, const_cast<uint64_t*>(d_offsets), 0,
                     d_hit_query, d_hit_target, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int bcu_join_dev(const bcu_index* ix, uint64_t n_q, const uint32_t* d_qgroup,
                            const uint32_t* d_qlow, const uint32_t* d_qhigh, uint64_t* d_offsets,
                            uint64_t pair_capacity, uint32_t* d_hit_query, uint32_t* d_hit_target,
                            uint64_t* d_total, uint32_t query_id_base, void* stream) {
  BCU_TRY(check_query_args("bcu_join_dev", ix, n_q, d_qlow, d_qhigh));
  if (!d_offsets || (pair_capacity && !d_hit_target)) {
    set_error("bcu_join_dev: NULL output pointer");
    return BCU_E_INVALID;
  }
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  return launch_join(ix, kModeFused, n_q, d_qgroup, d_qlow, d_qhigh, d_offsets, pair_capacity, d_hit_query,
                     d_hit_target, d_total, nullptr, query_id_base, static_cast<cudaStream_t>(stream));
}

extern "C" int bcu_join_filtered_dev(const bcu_index* ix, const bcu_filter* filter, uint64_t n_q,
                                     const uint32_t* d_qgroup, const uint32_t* d_qlow, const uint32_t* d_qhigh,
                                     const uint8_t* d_qstrand, uint64_t* d_offsets, uint64_t pair_capacity,
                                     uint32_t* d_hit_query, uint32_t* d_hit_target, uint64_t* d_total,
                                     uint32_t query_id_base, void* stream) {
  BCU_TRY(check_query_args("bcu_join_filtered_dev", ix, n_q, d_qlow, d_qhigh));
  if (!filter || !d_offsets || (pair_capacity && !d_hit_target)) {
    set_error("bcu_join_filtered_dev: NULL filter/output pointer");
    return BCU_E_INVALID;
  }
  if (filter->kind == BCU_FILTER_SV2NL_INV && filter->use_strand && n_q && !d_qstrand) {
    set_error("bcu_join_filtered_dev: the INV filter with use_strand needs d_qstrand");
    return BCU_E_INVALID;
  }
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  return launch_join(ix, kModeFused, n_q, d_qgroup, d_qlow, d_qhigh, d_offsets, pair_capacity, d_hit_query,
                     d_hit_target, d_total, nullptr, query_id_base, static_cast<cudaStream_t>(stream), nullptr,
                     filter, d_qstrand);
}

extern "C" int bcu_query_any_dev(const bcu_index* ix, uint64_t n_q, const uint32_t* d_qgroup,
                                 const uint32_t* d_qlow, const uint32_t* d_qhigh, uint8_t* d_any,
                                 void* stream) {
  BCU_TRY(check_query_args("bcu_query_any_dev", ix, n_q, d_qlow, d_qhigh));
  if (n_q && !d_any) { set_error("bcu_query_any_dev: d_any is NULL"); return BCU_E_INVALID; }
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  return launch_join(ix, kModeAny, n_q, d_qgroup, d_qlow, d_qhigh, nullptr, 0, nullptr, nullptr, nullptr,
                     d_any, 0, static_cast<cudaStream_t>(stream));
}

// ---- host-pointer entry points ---------------------------------------------------------------------
extern "C" int bcu_query_count(const bcu_index* ix, uint64_t n_q, const uint32_t* qgroup,
                               const uint32_t* qlow, const uint32_t* qhigh, uint64_t* offsets,
                               uint64_t* total) {
  BCU_TRY(check_query_args("bcu_query_count", ix, n_q, qlow, qhigh));
  if (!offsets) { set_error("bcu_query_count: offsets is NULL"); return BCU_E_INVALID; }
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  Staging st;
  BCU_TRY(st.init());
  uint32_t *d_g, *d_l, *d_h;
  uint64_t* d_off;
  BCU_TRY(st.upload(&d_g, qgroup, n_q));
  BCU_TRY(st.upload(&d_l, qlow, n_q));
  BCU_TRY(st.upload(&d_h, qhigh, n_q));
  BCU_TRY(st.alloc(&d_off, n_q + 1));
  BCU_TRY(launch_join(ix, kModeCount, n_q, d_g, d_l, d_h, d_off, 0, nullptr, nullptr, nullptr, nullptr, 0,
                      st.stream));
  BCU_CUDA(cudaMemcpyAsync(offsets, d_off, (n_q + 1) * 8, cudaMemcpyDeviceToHost, st.stream));
  BCU_CUDA(cudaStreamSynchronize(st.stream));
  if (total) *total = offsets[n_q];
  return BCU_OK;
}

extern "C" int bcu_query_scatter(const bcu_index* ix, uint64_t n_q, const uint32_t* qgroup,
                                 const uint32_t* qlow, const uint32_t* qhigh, const uint64_t* offsets,
                                 uint32_t* hit_query, uint32_t* hit_target) {
  BCU_TRY(check_query_args("bcu_query_scatter", ix, n_q, qlow, qhigh));
  if (!offsets) { set_error("bcu_query_scatter: offsets is NULL"); return BCU_E_INVALID; }
  const uint64_t total = offsets[n_q];
  if (total && (!hit_query || !hit_target)) {
    set_error("bcu_query_scatter: pair buffers are NULL");
    return BCU_E_INVALID;
  }
  if (n_q == 0 || total == 0) return BCU_OK;
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  Staging st;
  BCU_TRY(st.init());
  uint32_t *d_g, *d_l, *d_h, *d_hq, *d_ht;
  uint64_t* d_off;
  BCU_TRY(st.upload(&d_g, qgroup, n_q));
  BCU_TRY(st.upload(&d_l, qlow, n_q));
  BCU_TRY(st.upload(&d_h, qhigh, n_q));
  BCU_TRY(st.upload(&d_off, offsets, n_q + 1));
  BCU_TRY(st.alloc(&d_hq, total));
  BCU_TRY(st.alloc(&d_ht, total));
  BCU_TRY(launch_join(ix, kModeScatter, n_q, d_g, d_l, d_h, d_off, 0, d_hq, d_ht, nullptr, nullptr, 0,
                      st.stream));
  BCU_CUDA(cudaMemcpyAsync(hit_query, d_hq, total * 4, cudaMemcpyDeviceToHost, st.stream));
  BCU_CUDA(cudaMemcpyAsync(hit_target, d_ht, total * 4, cudaMemcpyDeviceToHost, st.stream));
  BCU_CUDA(cudaStreamSynchronize(st.stream));
  return BCU_OK;
}

extern "C" int bcu_query_any(const bcu_index* ix, uint64_t n_q, const uint32_t* qgroup,
                             const uint32_t* qlow, const uint32_t* qhigh, uint8_t* any) {
  BCU_TRY(check_query_args("bcu_query_any", ix, n_q, qlow, qhigh));
  if (n_q && !any) { set_error("bcu_query_any: any is NULL"); return BCU_E_INVALID; }
  if (n_q == 0) return BCU_OK;
  DeviceGuard guard(ix->device);
  BCU_TRY(guard_ok(guard, ix->device));
  Staging st;
  BCU_TRY(st.init());
  uint32_t *d_g, *d_l, *d_h;
  uint8_t* d_any;
  BCU_TRY(st.upload(&d_g, qgroup, n_q));
  BCU_TRY(st.upload(&d_l, qlow, n_q));
  BCU_TRY(st.upload(&d_h, qhigh, n_q));
  BCU_TRY(st.alloc(&d_any, n_q));
  BCU_TRY(launch_join(ix, kModeAny, n_q, d_g, d_l, d_h, nullptr, 0, nullptr, nullptr, nullptr, d_any, 0,
                      st.stream));
  BCU_CUDA(cudaMemcpyAsync(any, d_any, n_q, cudaMemcpyDeviceToHost, st.stream));
  BCU_CUDA(cudaStreamSynchronize(st.stream));
  return BCU_OK;
}
