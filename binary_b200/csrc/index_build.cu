// K2 -- construction of the flat augmented index that replaces the pointer-chasing tree.
//
// Reference being replaced: IntervalTree::insert_node_impl + max maintenance
// (interval_tree.hpp:230-260, 206-228, 262-278): a BST on `low` whose nodes carry the subtree max of
// `high`. Flat equivalent (NCList/AIList style):
//   rows sorted by (group, low, id)            <- K1 (radix_sort.cu), stable so ties keep id order
//   lowhigh[r] = {low, high}, id[r]            <- gather
//   runmax[r]  = max(high[group_begin..r])     <- segmented running max (scan.cu), the "max-end" array
//   dir[g][b]  = { first row with runmax >= b*W , first row with low >= (b+1)*W , {low, high} of that
//                first row },  W = 1 << shift: 16 bytes, ONE 128-bit load
// The directory turns both searches of a query (upper bound of q.high in `low`, lower bound of q.low in
// `runmax`) into one load when the query lies inside one bin (two otherwise) and brings the first
// candidate row along; the few rows of slack it admits are rejected by the exact predicate in the scan
// (join.cu), so results do not depend on W.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace bcu {

constexpr int kThreads = 256;

// keys[i] = group << 32 | low, vals[i] = i, and *varying |= key ^ key[0] (which bits differ at all)
__global__ void __launch_bounds__(kThreads)
    make_keys_kernel(const uint32_t* __restrict__ group, const uint32_t* __restrict__ low, uint64_t n,
                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                     unsigned long long* varying) {
  const uint64_t key0 = ((uint64_t)(group ? group[0] : 0u) << 32) | low[0];
  uint64_t diff = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t k = ((uint64_t)(group ? group[i] : 0u) << 32) | low[i];
    keys[i] = k;
    vals[i] = (uint32_t)i;
    diff |= k ^ key0;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) diff |= __shfl_xor_sync(0xffffffffu, (unsigned long long)diff, off);
  if ((threadIdx.x & 31) == 0 && diff) atomicOr(varying, (unsigned long long)diff);
}

// sorted rows -> {low, high} + id (+ a plain `high` column for the scan); group heads are appended to
// head_rows in arbitrary order (the host sorts them: there are few).
__global__ void __launch_bounds__(kThreads)
    gather_rows_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                       const uint32_t* __restrict__ high, uint64_t n, uint2* __restrict__ lowhigh,
                       uint32_t* __restrict__ ids, uint32_t* __restrict__ high_sorted,
                       uint32_t* __restrict__ head_rows, uint32_t* n_heads) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  uint64_t k = keys[r];
  uint32_t id = vals[r];
  uint32_t h = high[id];
  lowhigh[r] = make_uint2((uint32_t)k, h);
  ids[r] = id;
  high_sorted[r] = h;
  if (r == 0 || (uint32_t)(keys[r - 1] >> 32) != (uint32_t)(k >> 32)) {
    uint32_t slot = atomicAdd(n_heads, 1u);
    head_rows[slot] = (uint32_t)r;
  }
}

// per group: value, last row's low and runmax (= max high of the group) -> host picks the bin width
__global__ void group_probe_kernel(const uint32_t* __restrict__ head_rows, uint32_t n_groups, uint64_t n,
                                   const uint64_t* __restrict__ keys, const uint32_t* __restrict__ runmax,
                                   uint32_t* __restrict__ gval, uint32_t* __restrict__ cmax) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  uint32_t b = head_rows[g];
  uint64_t e = (g + 1 < n_groups) ? head_rows[g + 1] : n;
  gval[g] = (uint32_t)(keys[b] >> 32);
  uint32_t last_low = (uint32_t)keys[e - 1];
  uint32_t mh = runmax[e - 1];
  cmax[g] = last_low > mh ? last_low : mh;
}

// one thread per directory entry
__global__ void __launch_bounds__(kThreads)
    fill_directory_kernel(const GroupDesc* __restrict__ groups, uint32_t n_groups, uint32_t shift,
                          const uint2* __restrict__ lowhigh, const uint32_t* __restrict__ runmax,
                          DirEntry* __restrict__ dir, uint64_t n_bins) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_bins) return;
  // group owning entry e: last descriptor with bin_base <= e
  uint32_t lo = 0, hi = n_groups;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (groups[mid].bin_base <= e) lo = mid; else hi = mid;
  }
  const GroupDesc g = groups[lo];
  const uint64_t t0 = (e - g.bin_base) << shift;  // b * W       (<= cmax, fits 32 bits)
  const uint64_t t1 = t0 + (1ull << shift);       // (b + 1) * W (may exceed 32 bits in the last bin)
  uint32_t a = g.row_begin, b = g.row_end;
  while (a < b) {  // first row with runmax >= b*W
    uint32_t m = a + ((b - a) >> 1);
    if (runmax[m] < (uint32_t)t0) a = m + 1; else b = m;
  }
  const uint32_t lb = a;
  uint32_t ub = g.row_end;
  if (t1 <= 0xffffffffull) {
    a = g.row_begin; b = g.row_end;
    while (a < b) {  // first row with low >= (b+1)*W
      uint32_t m = a + ((b - a) >> 1);
      if (lowhigh[m].x < (uint32_t)t1) a = m + 1; else b = m;
    }
    ub = a;
  }
  DirEntry de;
  de.lb = lb;
  de.ub = ub;
  const uint2 r0 = (lb < g.row_end) ? lowhigh[lb] : make_uint2(0xffffffffu, 0u);
  de.low0 = r0.x;
  de.high0 = r0.y;
  dir[e] = de;
}

static double env_double(const char* name, double dflt) {
  const char* s = std::getenv(name);
  if (!s || !*s) return dflt;
  char* end = nullptr;
  double v = std::strtod(s, &end);
  return (end && end != s && v > 0) ? v : dflt;
}

static void free_index_members(bcu_index* ix) {
  cudaFree(ix->d_lowhigh);
  cudaFree(ix->d_id);
  cudaFree(ix->d_high);
  cudaFree(ix->d_runmax);
  cudaFree(ix->d_groups);
  cudaFree(ix->d_dir);
}

struct TempBuffers {  // freed on every exit path
  cudaStream_t stream;
  std::vector<void*> ptrs;
  explicit TempBuffers(cudaStream_t s) : stream(s) {}
  ~TempBuffers() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
  }
  template <class T> cudaError_t alloc(T** p, uint64_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, std::max<uint64_t>(count, 1) * sizeof(T), stream);
    if (e == cudaSuccess) ptrs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return e;
  }
};

// Keep stream-ordered scratch cached in the device pool (the default threshold of 0 hands it back to
// the driver at every synchronisation, which costs milliseconds per call).
static void keep_pool_warm(int device) {
  static std::atomic<unsigned long long> done{0};
  if (device < 0 || device >= 64 || (done.load() >> device) & 1ull) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  cudaGetLastError();
  done.fetch_or(1ull << device);
}

static int build_on_device(bcu_index* ix, uint64_t n, const uint32_t* d_group, const uint32_t* d_low,
                           const uint32_t* d_high, cudaStream_t stream) {
  ix->n = n;
  keep_pool_warm(ix->device);
  if (n == 0) return BCU_OK;
  TempBuffers tmp(stream);
  uint64_t *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b, *head_rows, *counters;
  BCU_CUDA(tmp.alloc(&keys_a, n));
  BCU_CUDA(tmp.alloc(&keys_b, n));
  BCU_CUDA(tmp.alloc(&vals_a, n));
  BCU_CUDA(tmp.alloc(&vals_b, n));
  BCU_CUDA(tmp.alloc(&head_rows, n));
  BCU_CUDA(tmp.alloc(&counters, 4));  // [0..1] varying bits (u64), [2] n_heads
  BCU_CUDA(cudaMemsetAsync(counters, 0, 16, stream));

  const unsigned grid_n = (unsigned)std::min<uint64_t>((n + kThreads - 1) / kThreads, 148ull * 16);
  make_keys_kernel<<<grid_n, kThreads, 0, stream>>>(d_group, d_low, n, keys_a, vals_a,
                                                   reinterpret_cast<unsigned long long*>(counters));
  BCU_LAUNCHED();
  uint64_t varying = 0;
  BCU_CUDA(cudaMemcpyAsync(&varying, counters, 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));

  uint64_t* keys;
  uint32_t* vals;
  BCU_TRY(radix_sort_pairs(keys_a, keys_b, vals_a, vals_b, n, varying, stream, &keys, &vals,
                           &ix->sort_passes));

  BCU_CUDA(cudaMalloc((void**)&ix->d_lowhigh, (n + 2) * sizeof(uint2)));  // +pad: join.cu reads row pairs
  BCU_CUDA(cudaMemsetAsync(ix->d_lowhigh + n, 0, 2 * sizeof(uint2), stream));
  BCU_CUDA(cudaMalloc((void**)&ix->d_id, (n + 4) * 4));  // +pad: 128-bit loads of 4 rows
  BCU_CUDA(cudaMalloc((void**)&ix->d_high, (n + 4) * 4));
  BCU_CUDA(cudaMemsetAsync(ix->d_id + n, 0, 16, stream));
  BCU_CUDA(cudaMemsetAsync(ix->d_high + n, 0, 16, stream));
  BCU_CUDA(cudaMalloc((void**)&ix->d_runmax, n * 4));
  ix->bytes += n * 20;
  const unsigned grid_rows = (unsigned)((n + kThreads - 1) / kThreads);
  gather_rows_kernel<<<grid_rows, kThreads, 0, stream>>>(keys, vals, d_high, n, ix->d_lowhigh, ix->d_id,
                                                        ix->d_high, head_rows, counters + 2);
  BCU_LAUNCHED();
  BCU_TRY(segmented_running_max(keys, ix->d_high, ix->d_runmax, n, stream));

  // ---- groups: sort the head rows on the host (few), probe value / max coordinate per group ----
  uint32_t n_groups = 0;
  BCU_CUDA(cudaMemcpyAsync(&n_groups, counters + 2, 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  std::vector<uint32_t> heads(n_groups);
  BCU_CUDA(cudaMemcpyAsync(heads.data(), head_rows, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  std::sort(heads.begin(), heads.end());
  BCU_CUDA(cudaMemcpyAsync(head_rows, heads.data(), (size_t)n_groups * 4, cudaMemcpyHostToDevice, stream));
  uint32_t *d_gval, *d_cmax;
  BCU_CUDA(tmp.alloc(&d_gval, n_groups));
  BCU_CUDA(tmp.alloc(&d_cmax, n_groups));
  group_probe_kernel<<<(n_groups + kThreads - 1) / kThreads, kThreads, 0, stream>>>(
      head_rows, n_groups, n, keys, ix->d_runmax, d_gval, d_cmax);
  BCU_LAUNCHED();
  std::vector<uint32_t> gval(n_groups), cmax(n_groups);
  BCU_CUDA(cudaMemcpyAsync(gval.data(), d_gval, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaMemcpyAsync(cmax.data(), d_cmax, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));

  // ---- bin width: smallest shift whose directory stays within ~bin_factor entries per target ----
  const double factor = env_double("BCU_BIN_FACTOR", 2.0);
  const uint64_t budget = std::max<uint64_t>((uint64_t)(factor * (double)n), 1024) + 2ull * n_groups;
  uint32_t shift = 0;
  uint64_t n_bins = 0;
  for (shift = 0; shift <= 31; ++shift) {
    n_bins = 0;
    for (uint32_t g = 0; g < n_groups; ++g) n_bins += ((uint64_t)cmax[g] >> shift) + 1;
    if (n_bins <= budget) break;
  }
  if (shift > 31) shift = 31;
  std::vector<GroupDesc> descs(n_groups);
  n_bins = 0;
  for (uint32_t g = 0; g < n_groups; ++g) {
    descs[g].gval = gval[g];
    descs[g].row_begin = heads[g];
    descs[g].row_end = (g + 1 < n_groups) ? heads[g + 1] : (uint32_t)n;
    descs[g].nb = (uint32_t)(((uint64_t)cmax[g] >> shift) + 1);
    descs[g].bin_base = n_bins;
    n_bins += (uint64_t)descs[g].nb;
  }
  ix->n_groups = n_groups;
  ix->max_gval = n_groups ? gval[n_groups - 1] : 0;
  ix->shift = shift;
  ix->n_bins = n_bins;
  BCU_CUDA(cudaMalloc((void**)&ix->d_groups, (size_t)n_groups * sizeof(GroupDesc)));
  BCU_CUDA(cudaMalloc((void**)&ix->d_dir, n_bins * sizeof(DirEntry)));
  ix->bytes += (uint64_t)n_groups * sizeof(GroupDesc) + n_bins * sizeof(DirEntry);
  BCU_CUDA(cudaMemcpyAsync(ix->d_groups, descs.data(), (size_t)n_groups * sizeof(GroupDesc),
                           cudaMemcpyHostToDevice, stream));
  fill_directory_kernel<<<(unsigned)((n_bins + kThreads - 1) / kThreads), kThreads, 0, stream>>>(
      ix->d_groups, n_groups, shift, ix->d_lowhigh, ix->d_runmax, ix->d_dir, n_bins);
  BCU_LAUNCHED();
  BCU_CUDA(cudaStreamSynchronize(stream));  // descs/heads are host temporaries
  return BCU_OK;
}

}  // namespace bcu

using namespace bcu;

extern "C" int bcu_index_build_dev(int device, uint64_t n_t, const uint32_t* d_group,
                                   const uint32_t* d_low, const uint32_t* d_high, void* stream,
                                   bcu_index** out) {
  if (!out) { set_error("bcu_index_build: out is NULL"); return BCU_E_INVALID; }
  *out = nullptr;
  if (n_t > 0x7fffffffull) { set_error("bcu_index_build: n_t exceeds 2^31-1"); return BCU_E_LIMIT; }
  if (n_t && (!d_low || !d_high)) { set_error("bcu_index_build: low/high are NULL"); return BCU_E_INVALID; }
  DeviceGuard guard(device);
  if (!guard.ok) { set_error("bcu_index_build: cannot select CUDA device %d", device); return BCU_E_CUDA; }
  bcu_index* ix = new (std::nothrow) bcu_index();
  if (!ix) { set_error("bcu_index_build: host allocation failed"); return BCU_E_NOMEM; }
  ix->device = device;
  int rc = build_on_device(ix, n_t, d_group, d_low, d_high, static_cast<cudaStream_t>(stream));
  if (rc != BCU_OK) {
    free_index_members(ix);
    delete ix;
    return rc;
  }
  *out = ix;
  return BCU_OK;
}

extern "C" int bcu_index_build(int device, uint64_t n_t, const uint32_t* group, const uint32_t* low,
                               const uint32_t* high, bcu_index** out) {
  if (!out) { set_error("bcu_index_build: out is NULL"); return BCU_E_INVALID; }
  *out = nullptr;
  if (n_t > 0x7fffffffull) { set_error("bcu_index_build: n_t exceeds 2^31-1"); return BCU_E_LIMIT; }
  if (n_t && (!low || !high)) { set_error("bcu_index_build: low/high are NULL"); return BCU_E_INVALID; }
  DeviceGuard guard(device);
  if (!guard.ok) { set_error("bcu_index_build: cannot select CUDA device %d", device); return BCU_E_CUDA; }
  uint32_t *d_group = nullptr, *d_low = nullptr, *d_high = nullptr;
  int rc = BCU_OK;
  auto upload = [&](uint32_t** d, const uint32_t* h) -> int {
    BCU_CUDA(cudaMalloc((void**)d, std::max<uint64_t>(n_t, 1) * 4));
    BCU_CUDA(cudaMemcpy(*d, h, n_t * 4, cudaMemcpyHostToDevice));
    return BCU_OK;
  };
  if (n_t) {
    if (group) rc = upload(&d_group, group);
    if (rc == BCU_OK) rc = upload(&d_low, low);
    if (rc == BCU_OK) rc = upload(&d_high, high);
  }
  if (rc == BCU_OK) rc = bcu_index_build_dev(device, n_t, d_group, d_low, d_high, nullptr, out);
  cudaFree(d_group);
  cudaFree(d_low);
  cudaFree(d_high);
  return rc;
}

extern "C" int bcu_index_free(bcu_index* ix) {
  if (!ix) return BCU_OK;
  DeviceGuard guard(ix->device);
  free_index_members(ix);
  delete ix;
  return BCU_OK;
}

extern "C" int bcu_index_size(const bcu_index* ix, uint64_t* n_t) {
  if (!ix || !n_t) { set_error("bcu_index_size: NULL argument"); return BCU_E_INVALID; }
  *n_t = ix->n;
  return BCU_OK;
}

extern "C" int bcu_index_get_info(const bcu_index* ix, bcu_index_info* info) {
  if (!ix || !info) { set_error("bcu_index_get_info: NULL argument"); return BCU_E_INVALID; }
  info->n_targets = ix->n;
  info->n_groups = ix->n_groups;
  info->n_components = 1;
  info->bin_shift = ix->shift;
  info->sort_passes = ix->sort_passes;
  info->n_bins = ix->n_bins;
  info->device_bytes = ix->bytes;
  info->device = ix->device;
  info->reserved = 0;
  return BCU_OK;
}
