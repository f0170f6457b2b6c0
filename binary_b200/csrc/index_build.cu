// K2 -- construction of the flat augmented index that replaces the pointer-chasing tree.
//
// Reference being replaced: IntervalTree::insert_node_impl + max maintenance
// (interval_tree.hpp:230-260, 206-228, 262-278): a BST on `low` whose nodes carry the subtree max of
// `high`. Flat equivalent (NCList/AIList style):
//   rows sorted by (class, group, low, id)     <- K1 (radix_sort.cu), stable so ties keep id order; `class`
//                                                 is the length class of the AIList-style decomposition
//                                                 (one class unless some targets are far longer than others)
//   lowhigh[r] = {low, high}, id[r]            <- gather
//   runmax[r]  = max(high[group_begin..r])     <- segmented running max (scan.cu), the "max-end" array
//   dir[g][b]  = { first row with runmax >= b*W , first row with low >= (b+1)*W , {low, high} of that
//                first row },  W = 1 << shift (one shift per length class): 16 bytes, ONE 128-bit load
//   hs[r], dirh[g][b]                          <- second sort on (segment, high): each segment's highs
//                                                 ascending + the first index with value >= b*W; with them a
//                                                 long candidate range is COUNTED by two rank lookups
//                                                 (join.cu count_by_ranks) when the segment has no inverted row
// and, after this file's build: bcu_index_image_size / export_dev / import_dev (one device blob per index).
// The directory turns both searches of a query (upper bound of q.high in `low`, lower bound of q.low in
// `runmax`) into one load when the query lies inside one bin (two otherwise) and brings the first
// candidate row along; the few rows of slack it admits are rejected by the exact predicate in the scan
// (join.cu), so results do not depend on W.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
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

// sorted rows -> {low, high}, id, plain `high` column and the segment key (= group value); segment heads
// are appended to head_rows in arbitrary order (the host sorts them: there are few).
__global__ void __launch_bounds__(kThreads)
    gather_rows_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                       const uint32_t* __restrict__ high, uint64_t n, uint2* __restrict__ lowhigh,
                       uint32_t* __restrict__ ids, uint32_t* __restrict__ high_sorted,
                       uint64_t* __restrict__ segkey, uint32_t* __restrict__ head_rows, uint32_t* n_heads,
                       uint32_t* max_len, unsigned long long* sum_len) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t len = 0;
  if (r < n) {
    uint64_t k = keys[r];
    uint32_t id = vals[r];
    uint32_t h = high[id];
    uint32_t l = (uint32_t)k;
    lowhigh[r] = make_uint2(l, h);
    ids[r] = id;
    high_sorted[r] = h;
    segkey[r] = k >> 32;
    len = h >= l ? h - l : 0u;  // inverted rows count as length 0
    if (r == 0 || (keys[r - 1] >> 32) != (k >> 32)) head_rows[atomicAdd(n_heads, 1u)] = (uint32_t)r;
  }
  unsigned long long total = len;
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    len = max(len, __shfl_xor_sync(0xffffffffu, len, off));
    total += __shfl_xor_sync(0xffffffffu, total, off);
  }
  if ((threadIdx.x & 31) == 0 && len) {
    atomicMax(max_len, len);
    atomicAdd(sum_len, total);
  }
}

// length class of every row (see build_on_device) as the key of a stable 2-bit partition
__global__ void __launch_bounds__(kThreads)
    length_class_kernel(const uint2* __restrict__ lowhigh, uint64_t n, uint32_t base_len, uint32_t n_class,
                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint2 t = lowhigh[r];
  const uint64_t len = t.y >= t.x ? (uint64_t)(t.y - t.x) : 0ull;
  uint32_t c = 0;
  uint64_t limit = base_len;  // class c holds lengths < base_len * 4^c (last class: everything else)
  while (c + 1 < n_class && len >= limit) { ++c; limit *= 4; }
  keys[r] = c;
  vals[r] = (uint32_t)r;
}

// rows permuted into (component, group, low) order
__global__ void __launch_bounds__(kThreads)
    permute_rows_kernel(const uint64_t* __restrict__ comp_sorted, const uint32_t* __restrict__ perm, uint64_t n,
                        const uint2* __restrict__ lowhigh_in, const uint32_t* __restrict__ ids_in,
                        const uint64_t* __restrict__ segkey_in, uint2* __restrict__ lowhigh,
                        uint32_t* __restrict__ ids, uint32_t* __restrict__ high, uint64_t* __restrict__ segkey,
                        uint32_t* __restrict__ head_rows, uint32_t* n_heads) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t src = perm[r];
  const uint2 t = lowhigh_in[src];
  lowhigh[r] = t;
  ids[r] = ids_in[src];
  high[r] = t.y;
  const uint64_t sk = (comp_sorted[r] << 32) | segkey_in[src];  // component in bits 32.., group below
  segkey[r] = sk;
  bool head = (r == 0);
  if (!head) head = ((comp_sorted[r - 1] << 32) | segkey_in[perm[r - 1]]) != sk;
  if (head) head_rows[atomicAdd(n_heads, 1u)] = (uint32_t)r;
}

// per segment: key, last row's low and runmax (= max high of the segment) -> host picks the bin width
__global__ void group_probe_kernel(const uint32_t* __restrict__ head_rows, uint32_t n_segs, uint64_t n,
                                   const uint64_t* __restrict__ segkey, const uint2* __restrict__ lowhigh,
                                   const uint32_t* __restrict__ runmax, uint64_t* __restrict__ seg_out,
                                   uint32_t* __restrict__ cmax) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_segs) return;
  uint32_t b = head_rows[g];
  uint64_t e = (g + 1 < n_segs) ? head_rows[g + 1] : n;
  seg_out[g] = segkey[b];
  uint32_t last_low = lowhigh[e - 1].x;
  uint32_t mh = runmax[e - 1];
  cmax[g] = last_low > mh ? last_low : mh;
}

// keys of the second sort: (segment ordinal << 32) | high; also clears `proper[seg]` for inverted rows
__global__ void __launch_bounds__(kThreads)
    high_keys_kernel(const uint2* __restrict__ lowhigh, uint64_t n, const uint32_t* __restrict__ head_rows,
                     uint32_t n_segs, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                     uint32_t* __restrict__ proper, unsigned long long* varying) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t diff = 0;
  if (r < n) {
    uint32_t lo = 0, hi = n_segs;  // segment of row r: last head <= r
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (head_rows[mid] <= r) lo = mid; else hi = mid;
    }
    const uint2 t = lowhigh[r];
    const uint64_t k = ((uint64_t)lo << 32) | t.y;
    keys[r] = k;
    vals[r] = (uint32_t)r;
    if (t.x > t.y) proper[lo] = 0u;
    diff = k ^ (uint64_t)lowhigh[0].y;  // key of row 0 has segment ordinal 0
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) diff |= __shfl_xor_sync(0xffffffffu, (unsigned long long)diff, off);
  if ((threadIdx.x & 31) == 0 && diff) atomicOr(varying, (unsigned long long)diff);
}

__global__ void __launch_bounds__(kThreads)
    unpack_high_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ hs) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) hs[r] = (uint32_t)keys[r];
}

// dirh[e] = first index of the segment's slice of hs with hs >= b*W (one thread per directory entry)
__global__ void __launch_bounds__(kThreads)
    fill_dirh_kernel(const GroupDesc* __restrict__ groups, uint32_t n_groups,
                     const uint32_t* __restrict__ hs, uint32_t* __restrict__ dirh, uint64_t n_bins) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_bins) return;
  uint32_t lo = 0, hi = n_groups;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (groups[mid].bin_base <= e) lo = mid; else hi = mid;
  }
  const GroupDesc g = groups[lo];
  const uint32_t t0 = (uint32_t)((e - g.bin_base) << g.shift);
  uint32_t a = g.row_begin, b = g.row_end;
  while (a < b) {
    uint32_t m = a + ((b - a) >> 1);
    if (hs[m] < t0) a = m + 1; else b = m;
  }
  dirh[e] = a;
}

// one thread per directory entry
__global__ void __launch_bounds__(kThreads)
    fill_directory_kernel(const GroupDesc* __restrict__ groups, uint32_t n_groups,
                          const uint2* __restrict__ lowhigh, const uint32_t* __restrict__ runmax,
                          DirEntry* __restrict__ dir, uint64_t n_bins) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_bins) return;
  // segment owning entry e: last descriptor with bin_base <= e
  uint32_t lo = 0, hi = n_groups;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (groups[mid].bin_base <= e) lo = mid; else hi = mid;
  }
  const GroupDesc g = groups[lo];
  const uint64_t t0 = (e - g.bin_base) << g.shift;  // b * W       (<= cmax, fits 32 bits)
  const uint64_t t1 = t0 + (1ull << g.shift);       // (b + 1) * W (may exceed 32 bits in the last bin)
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

__global__ void __launch_bounds__(kThreads)
    split_low_kernel(const uint2* __restrict__ lowhigh, uint64_t n_padded, uint32_t* __restrict__ low) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_padded) low[r] = lowhigh[r].x;
}

static double env_double(const char* name, double dflt) {
  const char* s = std::getenv(name);
  if (!s || !*s) return dflt;
  char* end = nullptr;
  double v = std::strtod(s, &end);
  return (end && end != s && v > 0) ? v : dflt;
}

static void free_bin_layout(bcu_index* ix) {
  cudaFreeAsync(ix->d_bn_desc, nullptr);
  cudaFreeAsync(ix->d_bn_groups, nullptr);
  cudaFreeAsync(ix->d_bn_cell2bin, nullptr);
  cudaFreeAsync(ix->d_bn_cellbits, nullptr);
  cudaFreeAsync(ix->d_bn_blob, nullptr);
  ix->d_bn_desc = nullptr; ix->d_bn_groups = nullptr; ix->d_bn_cell2bin = nullptr; ix->d_bn_blob = nullptr;
  ix->d_bn_cellbits = nullptr;
  ix->bn_bins = 0;
  cudaGetLastError();
}

static void free_index_members(bcu_index* ix) {
  // Index arrays live in the stream-ordered pool (see alloc_rows). *_dev queries are asynchronous and may
  // still be reading them on streams the legacy default stream does not wait for (cudaStreamNonBlocking,
  // e.g. PyTorch side streams): drain the device first, as a classic cudaFree would have done.
  cudaDeviceSynchronize();
  cudaFreeAsync(ix->d_lowhigh, nullptr);
  cudaFreeAsync(ix->d_id, nullptr);
  cudaFreeAsync(ix->d_high, nullptr);
  cudaFreeAsync(ix->d_runmax, nullptr);
  cudaFreeAsync(ix->d_groups, nullptr);
  cudaFreeAsync(ix->d_dir, nullptr);
  cudaFreeAsync(ix->d_hs, nullptr);
  cudaFreeAsync(ix->d_dirh, nullptr);
  cudaFreeAsync(ix->d_bn_low, nullptr);
  if (ix->bn_owns_rows) {
    cudaFreeAsync(ix->d_bn_high, nullptr);
    cudaFreeAsync(ix->d_bn_id, nullptr);
  }
  cudaFreeAsync(ix->d_bn_desc, nullptr);
  cudaFreeAsync(ix->d_bn_groups, nullptr);
  cudaFreeAsync(ix->d_bn_cell2bin, nullptr);
  cudaFreeAsync(ix->d_bn_cellbits, nullptr);
  cudaFreeAsync(ix->d_bn_blob, nullptr);
  cudaFreeAsync(ix->d_lc_seg, nullptr);
  cudaFreeAsync(ix->d_lc_off, nullptr);
  cudaFreeAsync(ix->d_lc_row0, nullptr);
  cudaFreeAsync(ix->d_lc_high, nullptr);
  cudaFreeAsync(ix->d_lc_id, nullptr);
  // the frees are ordered on the NULL stream: wait for them so that the pool can hand the memory to an allocation on
  // ANY stream right away (a rebuild on a non-blocking stream would otherwise grow the pool instead: ~8 ms extra)
  cudaStreamSynchronize(nullptr);
  cudaGetLastError();
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

static int alloc_rows(bcu_index* ix, uint64_t n, cudaStream_t stream) {
  // Pool allocations (cudaMallocAsync): a rebuild reuses cached device memory instead of paying the
  // driver's map/unmap cost of cudaMalloc/cudaFree (measured: 2 ms vs up to 300 ms per build).
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_lowhigh, (n + 4) * sizeof(uint2), stream));  // +pad: 4 rows per load
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_id, (n + 4) * 4, stream));                   // +pad: 4 rows per load
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_high, (n + 4) * 4, stream));
  BCU_CUDA(cudaMemsetAsync(ix->d_lowhigh + n, 0, 4 * sizeof(uint2), stream));
  BCU_CUDA(cudaMemsetAsync(ix->d_id + n, 0, 16, stream));
  BCU_CUDA(cudaMemsetAsync(ix->d_high + n, 0, 16, stream));
  return BCU_OK;
}

// heads (unordered, on device) -> sorted on the host -> segment keys and max coordinates
static int probe_segments(const bcu_index* ix, uint64_t n, uint32_t* head_rows, uint32_t* d_n_heads,
                          const uint64_t* segkey, TempBuffers& tmp, cudaStream_t stream,
                          std::vector<uint32_t>& heads, std::vector<uint64_t>& seg, std::vector<uint32_t>& cmax) {
  uint32_t n_segs = 0;
  BCU_CUDA(cudaMemcpyAsync(&n_segs, d_n_heads, 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  heads.resize(n_segs);
  BCU_CUDA(cudaMemcpyAsync(heads.data(), head_rows, (size_t)n_segs * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  std::sort(heads.begin(), heads.end());
  BCU_CUDA(cudaMemcpyAsync(head_rows, heads.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, stream));
  uint64_t* d_seg;
  uint32_t* d_cmax;
  BCU_CUDA(tmp.alloc(&d_seg, n_segs));
  BCU_CUDA(tmp.alloc(&d_cmax, n_segs));
  group_probe_kernel<<<(n_segs + kThreads - 1) / kThreads, kThreads, 0, stream>>>(
      head_rows, n_segs, n, segkey, ix->d_lowhigh, ix->d_runmax, d_seg, d_cmax);
  BCU_LAUNCHED();
  seg.resize(n_segs);
  cmax.resize(n_segs);
  BCU_CUDA(cudaMemcpyAsync(seg.data(), d_seg, (size_t)n_segs * 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaMemcpyAsync(cmax.data(), d_cmax, (size_t)n_segs * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  return BCU_OK;
}

static int build_on_device(bcu_index* ix, uint64_t n, const uint32_t* d_group, const uint32_t* d_low,
                           const uint32_t* d_high, cudaStream_t stream) {
  ix->n = n;
  keep_pool_warm(ix->device);
  if (n == 0) return BCU_OK;
  const bool trace = std::getenv("BCU_BUILD_TRACE") != nullptr;  // dev aid: phase timeline on stderr (adds syncs)
  const auto t_begin = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[bcu_index_build] %-28s %8.0f us\n", what,
            std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count());
  };
  TempBuffers tmp(stream);
  uint64_t *keys_a, *keys_b, *segkey;
  uint32_t *vals_a, *vals_b, *head_rows, *counters;
  BCU_CUDA(tmp.alloc(&keys_a, n));
  BCU_CUDA(tmp.alloc(&keys_b, n));
  BCU_CUDA(tmp.alloc(&segkey, n));
  BCU_CUDA(tmp.alloc(&vals_a, n));
  BCU_CUDA(tmp.alloc(&vals_b, n));
  BCU_CUDA(tmp.alloc(&head_rows, n));
  BCU_CUDA(tmp.alloc(&counters, 8));  // [0..1] varying bits (u64), [2] n_heads, [3] max length, [4] n_heads #2,
                                      // [6..7] sum of lengths (u64)
  BCU_CUDA(cudaMemsetAsync(counters, 0, 32, stream));

  mark("temporaries allocated");
  // ---- K1: sort by (group, low) --------------------------------------------------------------------
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

  mark("K1 sort");
  // ---- K2: rows, running max, segments ---------------------------------------------------------------
  BCU_TRY(alloc_rows(ix, n, stream));
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_runmax, n * 4, stream));
  ix->bytes += n * 20;
  const unsigned grid_rows = (unsigned)((n + kThreads - 1) / kThreads);
  gather_rows_kernel<<<grid_rows, kThreads, 0, stream>>>(keys, vals, d_high, n, ix->d_lowhigh, ix->d_id,
                                                        ix->d_high, segkey, head_rows, counters + 2,
                                                        counters + 3,
                                                        reinterpret_cast<unsigned long long*>(counters + 6));
  BCU_LAUNCHED();
  BCU_TRY(segmented_running_max(segkey, ix->d_high, ix->d_runmax, n, stream));
  std::vector<uint32_t> heads, cmax;
  std::vector<uint64_t> seg;
  BCU_TRY(probe_segments(ix, n, head_rows, counters + 2, segkey, tmp, stream, heads, seg, cmax));
  const uint32_t n_groups = (uint32_t)heads.size();
  std::vector<uint32_t> gval(n_groups);
  uint64_t span = 0;
  for (uint32_t g = 0; g < n_groups; ++g) { gval[g] = (uint32_t)seg[g]; span += (uint64_t)cmax[g] + 1; }
  // the (group, low)-sorted view the bin layout is built on (before any length-class permutation)
  std::vector<uint32_t> bn_begin(heads), bn_cmax(cmax);
  bn_begin.push_back((uint32_t)n);

  mark("rows, running max, segments");
  // ---- length classes (AIList-style decomposition) -------------------------------------------------
  // A target much longer than its neighbours drags the running max along and with it the first candidate
  // row of every later query. Targets are therefore split by length into up to 4 classes, each a 4x band
  // above base_len = max(16 * mean spacing of starts, 2 * mean length): inside a class the candidate window
  // of a query is at most ~4x the rows it can hit, and a handful of giant intervals ends up in a class of
  // its own. A broad but outlier-free length distribution (max < 4x that base) stays ONE class: splitting
  // it would only fragment the long-range scans. Classes become the top sort key (class, group, low); a
  // query probes every class slot.
  uint32_t max_len = 0;
  uint64_t sum_len = 0;
  BCU_CUDA(cudaMemcpyAsync(&max_len, counters + 3, 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaMemcpyAsync(&sum_len, counters + 6, 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  const double spacing = (double)span / (double)n, mean_len = (double)sum_len / (double)n;
  const double base_d = std::min(4.0e9, std::max({1.0, env_double("BCU_CLASS_SPACINGS", 16.0) * spacing,
                                                  env_double("BCU_CLASS_MEANS", 2.0) * mean_len}));
  const uint32_t base_len = (uint32_t)base_d;
  uint32_t n_class = 1;
  for (uint64_t limit = base_len; n_class < 4 && (uint64_t)max_len >= limit; limit *= 4) ++n_class;
  if (env_double("BCU_MAX_CLASSES", 4.0) < n_class) n_class = (uint32_t)env_double("BCU_MAX_CLASSES", 4.0);
  const uint32_t n_comp = n_class <= 1 ? 1u : (n_class == 2 ? 2u : 4u);  // slots: 1, 2 or 4 (see join.cu)

  if (n_comp > 1) {
    // stable partition of the sorted rows by class (one 2-bit radix pass), then permute
    length_class_kernel<<<grid_rows, kThreads, 0, stream>>>(ix->d_lowhigh, n, base_len, n_class, keys_a, vals_a);
    BCU_LAUNCHED();
    uint64_t* comp_sorted;
    uint32_t* perm;
    uint32_t passes = 0;
    BCU_TRY(radix_sort_pairs(keys_a, keys_b, vals_a, vals_b, n, 3ull, stream, &comp_sorted, &perm, &passes));
    ix->sort_passes += passes;
    uint64_t* segkey2;
    BCU_CUDA(tmp.alloc(&segkey2, n));
    bcu_index old = *ix;  // the (group, low)-ordered arrays become the source of the permutation
    ix->d_lowhigh = nullptr; ix->d_id = nullptr; ix->d_high = nullptr;
    int rc = alloc_rows(ix, n, stream);
    if (rc == BCU_OK) {
      permute_rows_kernel<<<grid_rows, kThreads, 0, stream>>>(comp_sorted, perm, n, old.d_lowhigh, old.d_id, segkey,
                                                            ix->d_lowhigh, ix->d_id, ix->d_high, segkey2,
                                                            head_rows, counters + 4);
      if (cudaGetLastError() != cudaSuccess) rc = BCU_E_CUDA;
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (rc == BCU_OK && cudaMallocAsync((void**)&ix->d_bn_low, (n + 4) * 4, stream) != cudaSuccess) rc = BCU_E_NOMEM;
    if (rc == BCU_OK) {  // the group-sorted rows stay, as plain columns, for the bin layout
      split_low_kernel<<<(unsigned)((n + 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(old.d_lowhigh, n + 4, ix->d_bn_low);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      ix->d_bn_high = old.d_high;
      ix->d_bn_id = old.d_id;
      ix->bn_owns_rows = true;
      ix->bytes += n * 12;
    }
    if (rc == BCU_OK && cudaStreamSynchronize(stream) != cudaSuccess) rc = BCU_E_CUDA;
    cudaFreeAsync(old.d_lowhigh, stream);
    if (rc != BCU_OK) {
      if (!ix->bn_owns_rows) { cudaFreeAsync(old.d_id, stream); cudaFreeAsync(old.d_high, stream); }
      set_error("index build: permuting rows into length classes failed");
      return rc;
    }
    segkey = segkey2;
    BCU_TRY(segmented_running_max(segkey, ix->d_high, ix->d_runmax, n, stream));
    BCU_TRY(probe_segments(ix, n, head_rows, counters + 4, segkey, tmp, stream, heads, seg, cmax));
  }
  const uint32_t n_segs = (uint32_t)heads.size();

  mark("length classes");
  // ---- second sorted view: each segment's `high` values ascending + "all rows proper" flags ----------
  // With every row proper (low <= high) and a proper query, {rows with high < q.low} is a subset of
  // {rows with low <= q.high}, so a range's hit count is a difference of two ranks and long ranges
  // need no counting scan (join.cu, probe). head_rows is sorted on the device at this point.
  uint32_t* d_proper;
  BCU_CUDA(tmp.alloc(&d_proper, n_segs));
  BCU_CUDA(cudaMemsetAsync(d_proper, 0xff, (size_t)n_segs * 4, stream));  // non-zero = proper
  BCU_CUDA(cudaMemsetAsync(counters, 0, 8, stream));
  high_keys_kernel<<<grid_rows, kThreads, 0, stream>>>(ix->d_lowhigh, n, head_rows, n_segs, keys_a, vals_a, d_proper,
                                                      reinterpret_cast<unsigned long long*>(counters));
  BCU_LAUNCHED();
  uint64_t varying_h = 0;
  std::vector<uint32_t> proper(n_segs);
  BCU_CUDA(cudaMemcpyAsync(&varying_h, counters, 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaMemcpyAsync(proper.data(), d_proper, (size_t)n_segs * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  {
    uint64_t* hk;
    uint32_t* hv;
    uint32_t passes = 0;
    BCU_TRY(radix_sort_pairs(keys_a, keys_b, vals_a, vals_b, n, varying_h, stream, &hk, &hv, &passes));
    ix->sort_passes += passes;
    BCU_CUDA(cudaMallocAsync((void**)&ix->d_hs, (n + 4) * 4, stream));
    ix->bytes += n * 4;
    unpack_high_kernel<<<grid_rows, kThreads, 0, stream>>>(hk, n, ix->d_hs);
    BCU_LAUNCHED();
  }

  mark("rank view (second sort)");
  // ---- bin width, per length class: the smallest shift whose directory stays within ~bin_factor entries
  // per row OF THAT CLASS. A sparse class (the few long intervals) thus gets wide bins and a directory small
  // enough to stay in L2, instead of one as large as the dense class's (bins follow coordinates, not rows).
  const double factor = env_double("BCU_BIN_FACTOR", 2.0);
  uint32_t class_shift[4] = {0, 0, 0, 0};
  for (uint32_t c = 0; c < n_comp; ++c) {
    uint64_t rows_c = 0, segs_c = 0;
    for (uint32_t s = 0; s < n_segs; ++s)
      if ((uint32_t)(seg[s] >> 32) == c) {
        rows_c += ((s + 1 < n_segs) ? heads[s + 1] : (uint32_t)n) - heads[s];
        ++segs_c;
      }
    const uint64_t budget = std::max<uint64_t>((uint64_t)(factor * (double)rows_c), 1024) + 2ull * segs_c;
    uint32_t sh = 0;
    for (sh = 0; sh < 31; ++sh) {
      uint64_t bins = 0;
      for (uint32_t s = 0; s < n_segs; ++s)
        if ((uint32_t)(seg[s] >> 32) == c) bins += ((uint64_t)cmax[s] >> sh) + 1;
      if (bins <= budget) break;
    }
    class_shift[c] = sh;
  }
  const uint32_t shift = class_shift[0];
  uint64_t n_bins = 0;
  // compact list of the segments that exist (directory fill) + the query-side table [component][group]
  std::vector<GroupDesc> segs(n_segs), table((size_t)n_comp * n_groups);
  for (uint32_t c = 0; c < n_comp; ++c)
    for (uint32_t g = 0; g < n_groups; ++g) {
      GroupDesc& d = table[(size_t)c * n_groups + g];
      d.gval = gval[g]; d.row_begin = d.row_end = 0; d.nb = 0; d.bin_base = 0; d.proper = 0; d.shift = class_shift[c];
    }
  n_bins = 0;
  for (uint32_t s = 0; s < n_segs; ++s) {
    GroupDesc& d = segs[s];
    d.gval = (uint32_t)seg[s];
    d.row_begin = heads[s];
    d.row_end = (s + 1 < n_segs) ? heads[s + 1] : (uint32_t)n;
    const uint32_t comp = (uint32_t)(seg[s] >> 32);
    d.shift = class_shift[comp];
    d.nb = (uint32_t)(((uint64_t)cmax[s] >> d.shift) + 1);
    d.bin_base = n_bins;
    d.proper = proper[s] ? 1u : 0u;
    n_bins += (uint64_t)d.nb;
    const uint32_t g = (uint32_t)(std::lower_bound(gval.begin(), gval.end(), d.gval) - gval.begin());
    table[(size_t)comp * n_groups + g] = d;
  }
  ix->n_groups = n_groups;
  ix->n_comp = n_comp;
  ix->class_base_len = base_len;
  ix->max_len = max_len;
  ix->max_gval = n_groups ? gval[n_groups - 1] : 0;
  ix->shift = shift;
  ix->shifts = class_shift[0] | (class_shift[1] << 8) | (class_shift[2] << 16) | (class_shift[3] << 24);
  ix->n_bins = n_bins;
  GroupDesc* d_segs;
  BCU_CUDA(tmp.alloc(&d_segs, n_segs));
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_groups, table.size() * sizeof(GroupDesc), stream));
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_dir, n_bins * sizeof(DirEntry), stream));
  ix->bytes += table.size() * sizeof(GroupDesc) + n_bins * sizeof(DirEntry);
  BCU_CUDA(cudaMemcpyAsync(d_segs, segs.data(), (size_t)n_segs * sizeof(GroupDesc), cudaMemcpyHostToDevice, stream));
  BCU_CUDA(cudaMemcpyAsync(ix->d_groups, table.data(), table.size() * sizeof(GroupDesc),
                           cudaMemcpyHostToDevice, stream));
  fill_directory_kernel<<<(unsigned)((n_bins + kThreads - 1) / kThreads), kThreads, 0, stream>>>(
      d_segs, n_segs, ix->d_lowhigh, ix->d_runmax, ix->d_dir, n_bins);
  BCU_LAUNCHED();
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_dirh, n_bins * 4, stream));
  ix->bytes += n_bins * 4;
  fill_dirh_kernel<<<(unsigned)((n_bins + kThreads - 1) / kThreads), kThreads, 0, stream>>>(
      d_segs, n_segs, ix->d_hs, ix->d_dirh, n_bins);
  BCU_LAUNCHED();
  BCU_CUDA(cudaStreamSynchronize(stream));  // host vectors are temporaries
  mark("directories");
  if (!ix->d_bn_low) {  // one class: the index rows already are in (group, low) order
    BCU_CUDA(cudaMallocAsync((void**)&ix->d_bn_low, (n + 4) * 4, stream));
    ix->bytes += n * 4;
    split_low_kernel<<<(unsigned)((n + 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(ix->d_lowhigh, n + 4, ix->d_bn_low);
    BCU_LAUNCHED();
    ix->d_bn_high = ix->d_high;
    ix->d_bn_id = ix->d_id;
  }
  if (n_comp == 1)  // the same bound an importer derives from the directory sizes: both sides build the same layout
    for (uint32_t g = 0; g < n_groups; ++g)
      bn_cmax[g] = (uint32_t)std::min<uint64_t>(((((uint64_t)bn_cmax[g] >> class_shift[0]) + 1) << class_shift[0]) - 1, 0xffffffffull);
  BCU_TRY(build_bin_layout(ix, gval.data(), bn_begin.data(), bn_cmax.data(), stream));
  if (ix->bn_bins == 0) {  // no layout: the (group, low)-sorted copies are not needed
    cudaFreeAsync(ix->d_bn_low, stream);
    ix->bytes -= n * 4;
    if (ix->bn_owns_rows) {
      cudaFreeAsync(ix->d_bn_high, stream);
      cudaFreeAsync(ix->d_bn_id, stream);
      ix->bytes -= n * 8;
    }
    ix->d_bn_low = ix->d_bn_high = ix->d_bn_id = nullptr;
    ix->bn_owns_rows = false;
  }
  mark("bin layout");
  BCU_TRY(build_long_lists(ix, stream));
  mark("stab lists");
  return BCU_OK;
}

// ---- stab lists (join.cu emit_long_kernel) ---------------------------------------------------------------
// A query whose candidate range [lb, ub) is long walks rows that start up to one maximal length before it; on
// densely covered data with a broad length law about half of them do not reach the query. Per segment and
// coordinate bin of width 2^lc_shift the index therefore keeps ONE contiguous block of {high, id} entries:
//   * the stab list of the bin: the rows with low < bin start <= high -- every hit that starts before the bin is
//     among them, and all but the few ending in the bin's first part are hits;
//   * then the bin's own rows (bin start <= low < next bin start) in row order.
// A query with q.low in the bin scans the block up to its exact upper bound (and, rarely, the rows of the following
// bins from the plain columns): the candidates shrink from (max_len + q.len) / spacing rows to
// coverage + (bin width / 2 + q.len) / spacing. Every segment has one extra, empty bin at its end so that
// row0[s + 1] is the end of bin s's rows for every real bin.
struct LcSeg { uint32_t row_begin, row_end, base, nb; };

__global__ void __launch_bounds__(kThreads)
    length_sum_kernel(const uint2* __restrict__ lowhigh, uint64_t n, unsigned long long* sum) {
  unsigned long long acc = 0;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
    const uint2 v = lowhigh[r];
    if (v.y > v.x) acc += v.y - v.x;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum, acc);
}

// FILL = false: slots[s] counts the stab-list entries of bin slot s; FILL = true: slots[s] is the slot's write
// cursor (starts at off[s]) and every row is also copied to its place among its own bin's rows, which end the block
template <bool FILL>
__global__ void __launch_bounds__(kThreads)
    lc_rows_kernel(const LcSeg* __restrict__ segs, uint32_t n_segs, const uint2* __restrict__ lowhigh,
                   const uint32_t* __restrict__ ids, uint64_t n, uint32_t shift, uint32_t* __restrict__ slots,
                   const uint32_t* __restrict__ off, const uint32_t* __restrict__ row0,
                   uint32_t* __restrict__ out_high, uint32_t* __restrict__ out_id) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint2 v = lowhigh[r];
  if (!FILL && v.y <= v.x) return;  // low < start <= high needs a proper row of length >= 1
  uint32_t lo = 0, hi = n_segs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (segs[mid].row_begin <= r) lo = mid; else hi = mid;
  }
  const LcSeg s = segs[lo];
  const uint32_t id = FILL ? ids[r] : 0u;
  if (FILL) {
    const uint32_t slot = s.base + min(v.x >> shift, s.nb - 2);  // (nb counts the empty end bin)
    const uint32_t p = off[slot + 1] - (row0[slot + 1] - (uint32_t)r);
    out_high[p] = v.y;
    out_id[p] = id;
    if (v.y <= v.x) return;
  }
  const uint32_t b0 = (v.x >> shift) + 1, b1 = min(v.y >> shift, s.nb - 2);
  for (uint32_t b = b0; b <= b1 && b >= b0; ++b) {
    const uint32_t p = atomicAdd(slots + s.base + b, 1u);
    if (FILL) { out_high[p] = v.y; out_id[p] = id; }
  }
}

// cnt[s] += rows of bin s (0 for the end bin of a segment)
__global__ void __launch_bounds__(kThreads)
    lc_block_size_kernel(const LcSeg* __restrict__ segs, uint32_t n_segs, const uint32_t* __restrict__ row0, uint64_t n_slots,
                         uint32_t* __restrict__ cnt) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  uint32_t lo = 0, hi = n_segs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (segs[mid].base <= i) lo = mid; else hi = mid;
  }
  if (i - segs[lo].base + 1 < segs[lo].nb) cnt[i] += row0[i + 1] - row0[i];
}

__global__ void __launch_bounds__(kThreads)
    lc_row0_kernel(const LcSeg* __restrict__ segs, uint32_t n_segs, const uint2* __restrict__ lowhigh, uint32_t shift,
                   uint64_t n_slots, uint32_t* __restrict__ row0) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  uint32_t lo = 0, hi = n_segs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (segs[mid].base <= i) lo = mid; else hi = mid;
  }
  const LcSeg s = segs[lo];
  const uint64_t x = (i - s.base) << shift;
  uint32_t a = s.row_begin, e = s.row_end;
  while (a < e) {
    const uint32_t m = a + ((e - a) >> 1);
    if ((uint64_t)lowhigh[m].x < x) a = m + 1; else e = m;
  }
  row0[i] = a;
}

// BCU_LONG_LISTS=0 never, =1 always (tests), default: when the summed target length is >= BCU_LONG_LISTS_COVER (16)
// times the summed coordinate extent of the groups, i.e. when nearly every query's candidate range is long.
int build_long_lists(bcu_index* ix, cudaStream_t stream) {
  const char* e = std::getenv("BCU_LONG_LISTS");
  const int mode = e ? std::atoi(e) : -1;
  if (mode == 0 || ix->n == 0 || ix->n_groups == 0) return BCU_OK;
  const uint64_t n = ix->n;
  const uint32_t n_slots = ix->n_comp * ix->n_groups;
  std::vector<GroupDesc> table(n_slots);
  TempBuffers tmp(stream);
  unsigned long long* d_sum;
  BCU_CUDA(tmp.alloc(&d_sum, 1));
  BCU_CUDA(cudaMemsetAsync(d_sum, 0, 8, stream));
  length_sum_kernel<<<(unsigned)std::min<uint64_t>((n + kThreads - 1) / kThreads, 4096), kThreads, 0, stream>>>(ix->d_lowhigh, n, d_sum);
  BCU_LAUNCHED();
  unsigned long long sum_len = 0;
  BCU_CUDA(cudaMemcpyAsync(&sum_len, d_sum, 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaMemcpyAsync(table.data(), ix->d_groups, table.size() * sizeof(GroupDesc), cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  // coordinate extent of each group as the directories bound it (an importer sees the same numbers)
  uint64_t extent = 0;
  for (uint32_t g = 0; g < ix->n_groups; ++g) {
    uint64_t x = 0;
    for (uint32_t c = 0; c < ix->n_comp; ++c) {
      const GroupDesc& d = table[(size_t)c * ix->n_groups + g];
      if (d.row_end > d.row_begin) x = std::max<uint64_t>(x, (uint64_t)d.nb << d.shift);
    }
    extent += x;
  }
  const char* ce = std::getenv("BCU_LONG_LISTS_COVER");
  const double need = ce ? std::atof(ce) : 16.0;
  if (mode != 1 && (double)sum_len < need * (double)extent) return BCU_OK;
  // bin width: a quarter of the mean length (about 4 list entries per row), and no row in more than 4096 lists
  const uint64_t mean_len = sum_len / n;
  uint32_t shift = 4;
  while (shift < 31 && ((1ull << shift) < mean_len / 4 || (1ull << shift) < (uint64_t)ix->max_len / 4096)) ++shift;
  if ((sum_len >> shift) + 2 * n > 0x7fffffffull) return BCU_OK;  // entries must stay 32-bit
  std::vector<LcSeg> segs;
  std::vector<uint2> seg_table(n_slots, make_uint2(0u, 0u));
  uint64_t n_bins = 0;
  for (uint32_t i = 0; i < n_slots; ++i) {  // class-major = ascending rows
    const GroupDesc& d = table[i];
    if (d.row_end <= d.row_begin) continue;
    const uint64_t cmax = ((uint64_t)d.nb << d.shift) - 1;
    LcSeg s;
    s.row_begin = d.row_begin; s.row_end = d.row_end;
    s.nb = (uint32_t)((cmax >> shift) + 2);  // + the empty end bin
    if (n_bins + s.nb > 0xffffffffull) return BCU_OK;
    s.base = (uint32_t)n_bins;
    seg_table[i] = make_uint2(s.base, s.nb);
    n_bins += s.nb;
    segs.push_back(s);
  }
  if (segs.empty() || (mode != 1 && n_bins > 4 * n + 4096ull * segs.size())) return BCU_OK;  // sparse extents
  LcSeg* d_segs;
  uint32_t* d_cnt;
  BCU_CUDA(tmp.alloc(&d_segs, segs.size()));
  BCU_CUDA(tmp.alloc(&d_cnt, n_bins + 1));
  BCU_CUDA(cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(LcSeg), cudaMemcpyHostToDevice, stream));
  BCU_CUDA(cudaMemsetAsync(d_cnt, 0, (n_bins + 1) * 4, stream));
  const unsigned row_grid = (unsigned)((n + kThreads - 1) / kThreads);
  const unsigned slot_grid = (unsigned)((n_bins + kThreads - 1) / kThreads);
  lc_rows_kernel<false><<<row_grid, kThreads, 0, stream>>>(d_segs, (uint32_t)segs.size(), ix->d_lowhigh, ix->d_id, n, shift,
                                                            d_cnt, nullptr, nullptr, nullptr, nullptr);
  BCU_LAUNCHED();
  uint32_t *d_off = nullptr, *d_row0 = nullptr;
  BCU_CUDA(cudaMallocAsync((void**)&d_off, (n_bins + 1) * 4, stream));
  if (cudaMallocAsync((void**)&d_row0, (n_bins + 1) * 4, stream) != cudaSuccess) {
    cudaFreeAsync(d_off, stream);
    set_error("stab lists: out of device memory");
    return BCU_E_NOMEM;
  }
  lc_row0_kernel<<<slot_grid, kThreads, 0, stream>>>(d_segs, (uint32_t)segs.size(), ix->d_lowhigh, shift, n_bins, d_row0);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  lc_block_size_kernel<<<slot_grid, kThreads, 0, stream>>>(d_segs, (uint32_t)segs.size(), d_row0, n_bins, d_cnt);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  int rc = exclusive_sum_u32(d_cnt, d_off, n_bins + 1, stream);
  uint32_t entries = 0;
  if (rc == BCU_OK && (cudaMemcpyAsync(&entries, d_off + n_bins, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                       cudaStreamSynchronize(stream) != cudaSuccess)) rc = BCU_E_CUDA;
  if (rc != BCU_OK || (mode != 1 && entries > 9 * n)) {  // (a few very long rows: the lists would outweigh the index)
    cudaFreeAsync(d_off, stream);
    cudaFreeAsync(d_row0, stream);
    if (rc != BCU_OK) set_error("stab lists: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  uint32_t *d_lh = nullptr, *d_li = nullptr;
  uint2* d_seg_table = nullptr;
  auto fail = [&](int code) {
    cudaFreeAsync(d_off, stream); cudaFreeAsync(d_row0, stream); cudaFreeAsync(d_lh, stream);
    cudaFreeAsync(d_li, stream); cudaFreeAsync(d_seg_table, stream);
    set_error("stab lists: %s", code == BCU_E_NOMEM ? "out of device memory" : cudaGetErrorString(cudaGetLastError()));
    return code;
  };
  if (cudaMallocAsync((void**)&d_lh, ((uint64_t)entries + 4) * 4, stream) != cudaSuccess ||
      cudaMallocAsync((void**)&d_li, ((uint64_t)entries + 4) * 4, stream) != cudaSuccess ||
      cudaMallocAsync((void**)&d_seg_table, n_slots * sizeof(uint2), stream) != cudaSuccess) return fail(BCU_E_NOMEM);
  if (cudaMemcpyAsync(d_cnt, d_off, (n_bins + 1) * 4, cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
      cudaMemcpyAsync(d_seg_table, seg_table.data(), n_slots * sizeof(uint2), cudaMemcpyHostToDevice, stream) != cudaSuccess)
    return fail(BCU_E_CUDA);
  lc_rows_kernel<true><<<row_grid, kThreads, 0, stream>>>(d_segs, (uint32_t)segs.size(), ix->d_lowhigh, ix->d_id, n, shift,
                                                           d_cnt, d_off, d_row0, d_lh, d_li);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return fail(BCU_E_CUDA);  // host vectors
  ix->lc_shift = shift;
  ix->lc_bins = n_bins;
  ix->lc_entries = entries;
  ix->d_lc_seg = d_seg_table;
  ix->d_lc_off = d_off;
  ix->d_lc_row0 = d_row0;
  ix->d_lc_high = d_lh;
  ix->d_lc_id = d_li;
  ix->bytes += (2 * n_bins + 2) * 4 + 2 * ((uint64_t)entries + 4) * 4 + n_slots * sizeof(uint2);
  return BCU_OK;
}

// ---- bin layout (binned_join.cu): coordinate tiles of the index that fit one CTA's shared memory ----------
// See common.cuh (BinDesc) for what the layout holds and why it answers a query exactly.

// For group g and cell boundary j (0..n_cells_g): the first row of the group with low >= j << shift.
// bound_base[g] = first boundary index of group g (n_cells_g + 1 boundaries per group).
__global__ void __launch_bounds__(kThreads)
    cell_rows_kernel(const uint32_t* __restrict__ group_begin, uint32_t n_groups, const uint32_t* __restrict__ bound_base,
                     uint32_t n_bounds, uint32_t shift, const uint32_t* __restrict__ low, uint32_t* __restrict__ start_row) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bounds) return;
  uint32_t lo = 0, hi = n_groups;  // group of boundary b: last g with bound_base[g] <= b
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (bound_base[mid] <= b) lo = mid; else hi = mid;
  }
  const uint64_t x = (uint64_t)(b - bound_base[lo]) << shift;
  uint32_t a = group_begin[lo], e = group_begin[lo + 1];
  while (a < e) {
    const uint32_t m = a + ((e - a) >> 1);
    if ((uint64_t)low[m] < x) a = m + 1; else e = m;
  }
  start_row[b] = a;
}

// sub_start of every (bin, sub-cell): tile-relative first own row with low >= start of the sub-cell.
// sub_base[b] = first global sub-cell slot of bin b (nsub_b + 1 slots per bin).
__global__ void __launch_bounds__(kThreads)
    sub_start_kernel(const BinDesc* __restrict__ desc, uint32_t n_bins, const uint32_t* __restrict__ sub_base,
                     uint32_t n_slots, const uint32_t* __restrict__ low, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  uint32_t lo = 0, hi = n_bins;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sub_base[mid] <= i) lo = mid; else hi = mid;
  }
  const BinDesc d = desc[lo];
  const uint64_t x = (uint64_t)d.x_begin + ((uint64_t)(i - sub_base[lo]) << d.ls);
  uint32_t a = d.row0 + d.lo, e = d.row0 + d.hi;
  while (a < e) {
    const uint32_t m = a + ((e - a) >> 1);
    if ((uint64_t)low[m] < x) a = m + 1; else e = m;
  }
  out[i] = a - d.row0;
}

// Coverage lists. One thread per row: every sub-cell start s with low < s <= high (over all the bins of the
// row's group that the row reaches) gets the row. FILL = false: count into cnt[slot]; FILL = true: the row's
// {high, id} is written at the slot's next free position (cursor[slot]) inside the bin's blob.
template <bool FILL>
__global__ void __launch_bounds__(kThreads)
    coverage_kernel(const uint32_t* __restrict__ low, const uint32_t* __restrict__ high, const uint32_t* __restrict__ ids,
                    uint32_t n, const uint32_t* __restrict__ group_begin, uint32_t n_groups,
                    const BinGroup* __restrict__ groups, const uint16_t* __restrict__ cell2bin, uint32_t cell_shift,
                    const BinDesc* __restrict__ desc, const uint32_t* __restrict__ sub_base, uint32_t* __restrict__ cnt,
                    const uint32_t* __restrict__ off, uint32_t* __restrict__ cursor, unsigned char* __restrict__ blob,
                    unsigned long long* total) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mine = 0;
  if (r < n) {
    const uint32_t l = low[r], h = high[r];
    if (h > l) {  // covers the sub-cell starts in (l, h]
      uint32_t lo = 0, hi = n_groups;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (group_begin[mid] <= r) lo = mid; else hi = mid;
      }
      const BinGroup g = groups[lo];
      const uint32_t b_first = cell2bin[g.cell_base + (l >> cell_shift)];
      const uint32_t b_last = cell2bin[g.cell_base + min(h >> cell_shift, g.n_cells - 1u)];
      for (uint32_t b = b_first; b <= b_last; ++b) {  // bins of one group are numbered consecutively
        const BinDesc d = desc[b];
        const uint32_t g_from = l < d.x_begin ? 0u : ((l - d.x_begin) >> d.ls) + 1u;
        if (h < d.x_begin) break;
        const uint32_t g_to = min((h - d.x_begin) >> d.ls, d.nsub - 1u);
        for (uint32_t s = g_from; s <= g_to && g_from < d.nsub; ++s) {
          const uint32_t slot = sub_base[b] + s;
          if (!FILL) {
            atomicAdd(&cnt[slot], 1u);
            ++mine;
          } else {
            const uint32_t at = off[slot] - off[sub_base[b]] + atomicAdd(&cursor[slot], 1u);
            const uint32_t tables = ((d.nsub + 1 + 7) & ~7u) * 4u;  // sub_start + cov_rel, u16 each
            uint32_t* cov_high = reinterpret_cast<uint32_t*>(blob + d.blob + tables);
            cov_high[at] = h;
            cov_high[d.n_cov + at] = ids[r];
          }
        }
      }
    }
  }
  if (!FILL) {
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
  }
}

// the two u16 tables at the head of every bin's blob
__global__ void __launch_bounds__(kThreads)
    blob_tables_kernel(const BinDesc* __restrict__ desc, uint32_t n_bins, const uint32_t* __restrict__ sub_base,
                       uint32_t n_slots, const uint32_t* __restrict__ sub_start, const uint32_t* __restrict__ off,
                       unsigned char* __restrict__ blob) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  uint32_t lo = 0, hi = n_bins;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sub_base[mid] <= i) lo = mid; else hi = mid;
  }
  const BinDesc d = desc[lo];
  const uint32_t s = i - sub_base[lo];
  uint16_t* tab = reinterpret_cast<uint16_t*>(blob + d.blob);
  tab[s] = (uint16_t)sub_start[i];
  tab[((d.nsub + 1 + 7) & ~7u) + s] = (uint16_t)(off[i] - off[sub_base[lo]]);
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n,
                                  uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

// One attempt with a given row target per bin. *too_big is set when a bin's rows + blob exceed the tile: the
// caller retries with a smaller target.
static int bin_layout_attempt(bcu_index* ix, const uint32_t* group_gval, const uint32_t* group_begin_h,
                              const uint32_t* group_cmax, cudaStream_t stream, uint32_t rows_target, double* too_big) {
  *too_big = 0.0;  // on return: by what factor the largest bin exceeds the tile (0 = every bin fits)
  const uint32_t G = ix->n_groups, n = (uint32_t)ix->n;
  // cell width: the routing table over all groups must stay within kBinMaxCells entries
  uint32_t shift = 0;
  for (; shift < 32; ++shift) {
    uint64_t cells = 0;
    for (uint32_t g = 0; g < G; ++g) cells += ((uint64_t)group_cmax[g] >> shift) + 1;
    if (cells <= kBinMaxCells) break;
  }
  if (shift >= 32) return BCU_OK;
  std::vector<BinGroup> groups(G);
  std::vector<uint32_t> bound_base(G + 1);
  uint32_t n_cells = 0;
  for (uint32_t g = 0; g < G; ++g) {
    groups[g].gval = group_gval[g];
    groups[g].cell_base = n_cells;
    groups[g].n_cells = (uint32_t)(((uint64_t)group_cmax[g] >> shift) + 1);
    groups[g].pad = 0;
    bound_base[g] = n_cells + g;
    n_cells += groups[g].n_cells;
  }
  bound_base[G] = n_cells + G;
  const uint32_t n_bounds = n_cells + G;

  TempBuffers tmp(stream);
  uint32_t *d_bound_base, *d_start, *d_group_begin;
  BCU_CUDA(tmp.alloc(&d_bound_base, G + 1));
  BCU_CUDA(tmp.alloc(&d_group_begin, G + 1));
  BCU_CUDA(tmp.alloc(&d_start, n_bounds));
  BCU_CUDA(cudaMemcpyAsync(d_bound_base, bound_base.data(), (G + 1) * 4, cudaMemcpyHostToDevice, stream));
  BCU_CUDA(cudaMemcpyAsync(d_group_begin, group_begin_h, (G + 1) * 4, cudaMemcpyHostToDevice, stream));
  cell_rows_kernel<<<(n_bounds + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_group_begin, G, d_bound_base, n_bounds,
                                                                               shift, ix->d_bn_low, d_start);
  BCU_LAUNCHED();
  std::vector<uint32_t> start(n_bounds);
  BCU_CUDA(cudaMemcpyAsync(start.data(), d_start, (size_t)n_bounds * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));

  // greedy: extend a bin cell by cell while it stays within the row target. The coverage lists of a bin are not
  // known yet; the target leaves them about a third of the tile (verified below).
  const uint32_t sub_rows = (uint32_t)std::max(1.0, env_double("BCU_BIN_SUBROWS", 6.0));  // own rows per sub-cell
  std::vector<BinDesc> bins;
  std::vector<uint16_t> cell2bin(n_cells, (uint16_t)kBinNull);
  std::vector<uint32_t> sub_base;
  uint32_t n_slots = 0;
  for (uint32_t g = 0; g < G; ++g) {
    const uint32_t nc = groups[g].n_cells;
    const uint32_t* st = start.data() + bound_base[g];
    for (uint32_t j0 = 0; j0 < nc;) {
      uint32_t j1 = j0 + 1;
      if (st[j1] - st[j0] + 8 > 12288) return BCU_OK;  // one cell alone exceeds any tile: not eligible
      while (j1 < nc && st[j1 + 1] - st[j0] <= rows_target) ++j1;
      if (bins.size() >= kBinMaxBins) return BCU_OK;
      BinDesc b;
      std::memset(&b, 0, sizeof(b));
      b.group = g;
      b.x_begin = (uint32_t)((uint64_t)j0 << shift);
      const bool last = j1 == nc;
      b.x_end = last ? 0u : (uint32_t)((uint64_t)j1 << shift);
      const uint64_t span = (last ? (uint64_t)group_cmax[g] + 1 : ((uint64_t)j1 << shift)) - b.x_begin;  // >= 1
      const uint32_t own = st[j1] - st[j0];
      b.row0 = st[j0] & ~3u;
      b.n_copy = own ? (st[j1] - b.row0 + 3) / 4 * 4 : 0u;
      b.lo = st[j0] - b.row0;
      b.hi = st[j1] - b.row0;
      const uint64_t want_sub = std::max<uint64_t>(1, own / sub_rows);
      uint32_t ls = 0;
      while (((span - 1) >> ls) + 1 > want_sub) ++ls;
      b.ls = ls;
      b.nsub = (uint32_t)(((span - 1) >> ls) + 1);
      for (uint32_t j = j0; j < j1; ++j) cell2bin[groups[g].cell_base + j] = (uint16_t)bins.size();
      sub_base.push_back(n_slots);
      n_slots += b.nsub + 1;
      bins.push_back(b);
      j0 = j1;
    }
  }
  const uint32_t K = (uint32_t)bins.size();
  if (K == 0) return BCU_OK;
  sub_base.push_back(n_slots);

  // device copies of the routing tables (kept by the index) and of the provisional descriptors
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_bn_desc, (size_t)K * sizeof(BinDesc), stream));
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_bn_groups, (size_t)G * sizeof(BinGroup), stream));
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_bn_cell2bin, (size_t)n_cells * 2, stream));
  BCU_CUDA(cudaMemcpyAsync(ix->d_bn_desc, bins.data(), (size_t)K * sizeof(BinDesc), cudaMemcpyHostToDevice, stream));
  BCU_CUDA(cudaMemcpyAsync(ix->d_bn_groups, groups.data(), (size_t)G * sizeof(BinGroup), cudaMemcpyHostToDevice, stream));
  BCU_CUDA(cudaMemcpyAsync(ix->d_bn_cell2bin, cell2bin.data(), (size_t)n_cells * 2, cudaMemcpyHostToDevice, stream));
  // the same map as a rank structure (bins are numbered in cell order): 1/10 of the bytes, which is what lets
  // two CTAs of the routing kernel share an SM
  const uint32_t n_words = (n_cells + 31) / 32;
  std::vector<uint32_t> cellbits(2 * (size_t)n_words, 0u);
  for (uint32_t i = 0; i < n_cells; ++i)
    if (i == 0 || cell2bin[i] != cell2bin[i - 1]) cellbits[i >> 5] |= 1u << (i & 31);
  for (uint32_t w = 1; w < n_words; ++w) cellbits[n_words + w] = cellbits[n_words + w - 1] + (uint32_t)__builtin_popcount(cellbits[w - 1]);
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_bn_cellbits, cellbits.size() * 4, stream));
  BCU_CUDA(cudaMemcpyAsync(ix->d_bn_cellbits, cellbits.data(), cellbits.size() * 4, cudaMemcpyHostToDevice, stream));
  uint32_t *d_sub_base, *d_sub_start, *d_cnt, *d_off, *d_bin_off;
  unsigned long long* d_total;
  BCU_CUDA(tmp.alloc(&d_sub_base, K + 1));
  BCU_CUDA(tmp.alloc(&d_sub_start, n_slots));
  BCU_CUDA(tmp.alloc(&d_cnt, (uint64_t)n_slots + 4));
  BCU_CUDA(tmp.alloc(&d_off, (uint64_t)n_slots + 4));
  BCU_CUDA(tmp.alloc(&d_bin_off, K + 1));
  BCU_CUDA(tmp.alloc(&d_total, 1));
  BCU_CUDA(cudaMemcpyAsync(d_sub_base, sub_base.data(), (size_t)(K + 1) * 4, cudaMemcpyHostToDevice, stream));
  BCU_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)n_slots + 4) * 4, stream));
  BCU_CUDA(cudaMemsetAsync(d_total, 0, 8, stream));
  sub_start_kernel<<<(n_slots + kThreads - 1) / kThreads, kThreads, 0, stream>>>(ix->d_bn_desc, K, d_sub_base, n_slots,
                                                                             ix->d_bn_low, d_sub_start);
  BCU_LAUNCHED();
  const unsigned grid_rows = (n + kThreads - 1) / kThreads;
  coverage_kernel<false><<<grid_rows, kThreads, 0, stream>>>(ix->d_bn_low, ix->d_bn_high, ix->d_bn_id, n, d_group_begin, G,
                                                            ix->d_bn_groups, ix->d_bn_cell2bin, shift, ix->d_bn_desc,
                                                            d_sub_base, d_cnt, nullptr, nullptr, nullptr, d_total);
  BCU_LAUNCHED();
  // one more slot than there are (its exclusive sum is the grand total), so that every bin's end is an entry of off
  BCU_TRY(exclusive_sum_u32(d_cnt, d_off, (uint64_t)n_slots + 1, stream));
  gather_u32_kernel<<<(K + 1 + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_off, d_sub_base, K + 1, d_bin_off);
  BCU_LAUNCHED();
  std::vector<uint32_t> bin_off(K + 1);
  unsigned long long total = 0;
  BCU_CUDA(cudaMemcpyAsync(bin_off.data(), d_bin_off, (size_t)(K + 1) * 4, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  // Eligibility: every bin's rows + blob fit the tile, and the lists stay a small multiple of the rows (long targets
  // cost one entry per sub-cell they cover: a set dominated by them is better served by the general path)
  if ((double)total > env_double("BCU_BINNED_COVER", 4.0) * (double)n + 65536.0) {
    free_bin_layout(ix);
    return BCU_OK;
  }
  uint64_t blob_bytes = 0;
  for (uint32_t b = 0; b < K; ++b) {
    BinDesc& d = bins[b];
    d.n_cov = (bin_off[b + 1] - bin_off[b] + 3) & ~3u;
    d.blob = blob_bytes;
    d.blob_bytes = ((d.nsub + 1 + 7) & ~7u) * 4u + d.n_cov * 8u;
    blob_bytes += d.blob_bytes;
    const double over = std::max((double)d.n_cov / 65535.0, ((double)d.n_copy * 12.0 + d.blob_bytes) / (double)kBinTileBytes);
    if (over > 1.0) *too_big = std::max(*too_big, over);
  }
  if (*too_big > 0.0) {
    free_bin_layout(ix);
    return BCU_OK;
  }
  uint32_t* d_cursor;
  BCU_CUDA(tmp.alloc(&d_cursor, (uint64_t)n_slots + 4));
  BCU_CUDA(cudaMemsetAsync(d_cursor, 0, ((size_t)n_slots + 4) * 4, stream));
  BCU_CUDA(cudaMallocAsync((void**)&ix->d_bn_blob, blob_bytes + 16, stream));
  BCU_CUDA(cudaMemsetAsync(ix->d_bn_blob, 0, blob_bytes + 16, stream));
  BCU_CUDA(cudaMemcpyAsync(ix->d_bn_desc, bins.data(), (size_t)K * sizeof(BinDesc), cudaMemcpyHostToDevice, stream));
  blob_tables_kernel<<<(n_slots + kThreads - 1) / kThreads, kThreads, 0, stream>>>(ix->d_bn_desc, K, d_sub_base, n_slots,
                                                                               d_sub_start, d_off, ix->d_bn_blob);
  BCU_LAUNCHED();
  coverage_kernel<true><<<grid_rows, kThreads, 0, stream>>>(ix->d_bn_low, ix->d_bn_high, ix->d_bn_id, n, d_group_begin, G,
                                                           ix->d_bn_groups, ix->d_bn_cell2bin, shift, ix->d_bn_desc,
                                                           d_sub_base, nullptr, d_off, d_cursor, ix->d_bn_blob, nullptr);
  BCU_LAUNCHED();
  BCU_CUDA(cudaStreamSynchronize(stream));  // host vectors are temporaries
  ix->bytes += (uint64_t)K * sizeof(BinDesc) + (size_t)G * sizeof(BinGroup) + (size_t)n_cells * 2 + blob_bytes;
  ix->bn_bins = K;
  ix->bn_cells = n_cells;
  ix->bn_cell_shift = shift;
  return BCU_OK;
}

int build_bin_layout(bcu_index* ix, const uint32_t* group_gval, const uint32_t* group_begin_h,
                     const uint32_t* group_cmax, cudaStream_t stream) {
  ix->bn_bins = 0;
  const char* off_env = std::getenv("BCU_BINNED");
  if (ix->n == 0 || ix->n_groups == 0 || ix->n_groups > (uint32_t)kMaxSmemGroups || (off_env && off_env[0] == '0'))
    return BCU_OK;
  // The binned path only pays for indexes far beyond L2 (launch_join_binned: >= 96 MB): smaller ones skip the
  // layout and its build time (BCU_BINNED=1 builds it regardless -- tests, experiments)
  const bool forced = off_env && off_env[0] == '1';
  if (!forced && (double)ix->n < env_double("BCU_BINNED_MIN_TARGETS", 2.0e6)) return BCU_OK;
  // the coverage lists of a bin are only known after they are counted: start from a target that leaves them a
  // third of the tile and shrink it while some bin does not fit
  uint32_t rows_target = (uint32_t)std::max(16.0, std::min(12288.0, env_double("BCU_BIN_ROWS", 8192.0)));
  for (int attempt = 0; attempt < 5; ++attempt) {
    double over = 0.0;
    BCU_TRY(bin_layout_attempt(ix, group_gval, group_begin_h, group_cmax, stream, rows_target, &over));
    if (over == 0.0) break;
    rows_target = (uint32_t)((double)rows_target / over * 0.96);  // the fullest bin decides, with a little room
    if (rows_target < 16) break;
  }
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
  DeviceGuard guard(ix->device);  // (if the device cannot be selected the frees fail harmlessly)
  free_index_members(ix);
  delete ix;
  return BCU_OK;
}

// ---- index image: every array of an index in ONE device blob -------------------------------------
// so that an index built on one GPU can be shipped to the others (NCCL broadcast or a peer copy over
// NVLink) instead of being rebuilt there. Layout: 256-byte header, then the arrays,
// each at a 256-byte boundary. `runmax` is a build-time array and is not part of the image.
namespace {
constexpr uint64_t kImageMagic = 0x3258444955434942ull;  // "BICUIDX2"
constexpr uint64_t kImageAlign = 256;
enum { kImgLowHigh, kImgHigh, kImgId, kImgHs, kImgGroups, kImgDir, kImgDirh, kImgArrays };
struct ImageHeader {
  uint64_t magic, total_bytes, n, n_bins, bytes;
  uint32_t n_groups, n_comp, class_base_len, shift, max_gval, sort_passes, shifts, max_len;
  uint64_t offset[kImgArrays], size[kImgArrays];
  uint8_t pad[256 - 5 * 8 - 8 * 4 - 2 * 8 * kImgArrays];
};
static_assert(sizeof(ImageHeader) == 256, "image header is one 256-byte block");

void image_layout(const bcu_index* ix, ImageHeader* h) {
  std::memset(h, 0, sizeof(*h));
  h->magic = kImageMagic;
  h->n = ix->n; h->n_bins = ix->n_bins; h->bytes = ix->bytes;
  h->n_groups = ix->n_groups; h->n_comp = ix->n_comp; h->class_base_len = ix->class_base_len;
  h->shift = ix->shift; h->max_gval = ix->max_gval; h->sort_passes = ix->sort_passes; h->shifts = ix->shifts;
  h->max_len = ix->max_len;
  const uint64_t rows = ix->n ? ix->n + 4 : 0;  // the padded row arrays travel with their padding
  h->size[kImgLowHigh] = rows * sizeof(uint2);
  h->size[kImgHigh] = h->size[kImgId] = h->size[kImgHs] = rows * 4;
  h->size[kImgGroups] = ix->n ? (uint64_t)ix->n_comp * ix->n_groups * sizeof(GroupDesc) : 0;
  h->size[kImgDir] = ix->n_bins * sizeof(DirEntry);
  h->size[kImgDirh] = ix->n_bins * 4;
  uint64_t at = sizeof(ImageHeader);
  for (int k = 0; k < kImgArrays; ++k) {
    h->offset[k] = at;
    at += (h->size[k] + kImageAlign - 1) / kImageAlign * kImageAlign;
  }
  h->total_bytes = at;
}
void** image_slots(bcu_index* ix, void** slot) {
  slot[kImgLowHigh] = &ix->d_lowhigh; slot[kImgHigh] = &ix->d_high; slot[kImgId] = &ix->d_id;
  slot[kImgHs] = &ix->d_hs; slot[kImgGroups] = &ix->d_groups; slot[kImgDir] = &ix->d_dir;
  slot[kImgDirh] = &ix->d_dirh;
  return slot;
}
}  // namespace

extern "C" int bcu_index_image_size(const bcu_index* ix, uint64_t* bytes) {
  if (!ix || !bytes) { set_error("bcu_index_image_size: NULL argument"); return BCU_E_INVALID; }
  ImageHeader h;
  image_layout(ix, &h);
  *bytes = h.total_bytes;
  return BCU_OK;
}

extern "C" int bcu_index_export_dev(const bcu_index* ix, void* d_image, uint64_t bytes, void* stream_) {
  if (!ix || !d_image) { set_error("bcu_index_export_dev: NULL argument"); return BCU_E_INVALID; }
  ImageHeader h;
  image_layout(ix, &h);
  if (bytes < h.total_bytes) {
    set_error("bcu_index_export_dev: image buffer of %llu bytes, %llu needed", (unsigned long long)bytes,
              (unsigned long long)h.total_bytes);
    return BCU_E_CAPACITY;
  }
  DeviceGuard guard(ix->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  void* slot[kImgArrays];
  image_slots(const_cast<bcu_index*>(ix), slot);
  char* img = static_cast<char*>(d_image);
  BCU_CUDA(cudaMemcpyAsync(img, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
  for (int k = 0; k < kImgArrays; ++k)
    if (h.size[k])
      BCU_CUDA(cudaMemcpyAsync(img + h.offset[k], *static_cast<void**>(slot[k]), h.size[k],
                               cudaMemcpyDeviceToDevice, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));  // the header is a stack temporary
  return BCU_OK;
}

extern "C" int bcu_index_import_dev(int device, const void* d_image, uint64_t bytes, void* stream_,
                                    bcu_index** out) {
  if (!out) { set_error("bcu_index_import_dev: out is NULL"); return BCU_E_INVALID; }
  *out = nullptr;
  if (!d_image || bytes < sizeof(ImageHeader)) { set_error("bcu_index_import_dev: no image"); return BCU_E_INVALID; }
  DeviceGuard guard(device);
  if (!guard.ok) { set_error("bcu_index_import_dev: cannot select CUDA device %d", device); return BCU_E_CUDA; }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ImageHeader h;
  BCU_CUDA(cudaMemcpyAsync(&h, d_image, sizeof(h), cudaMemcpyDeviceToHost, stream));
  BCU_CUDA(cudaStreamSynchronize(stream));
  bcu_index probe;  // recompute the layout from the header's scalars: the offsets must agree
  probe.n = h.n; probe.n_bins = h.n_bins; probe.n_groups = h.n_groups; probe.n_comp = h.n_comp;
  ImageHeader want;
  image_layout(&probe, &want);
  if (h.magic != kImageMagic || h.total_bytes != want.total_bytes || bytes < h.total_bytes ||
      std::memcmp(h.offset, want.offset, sizeof(h.offset)) || std::memcmp(h.size, want.size, sizeof(h.size)) ||
      h.n > 0x7fffffffull || (h.n_comp != 1 && h.n_comp != 2 && h.n_comp != 4)) {
    set_error("bcu_index_import_dev: not an index image (or truncated)");
    return BCU_E_INVALID;
  }
  bcu_index* ix = new (std::nothrow) bcu_index();
  if (!ix) { set_error("out of host memory"); return BCU_E_NOMEM; }
  ix->device = device;
  ix->n = h.n; ix->n_bins = h.n_bins; ix->bytes = h.bytes;
  ix->n_groups = h.n_groups; ix->n_comp = h.n_comp; ix->class_base_len = h.class_base_len;
  ix->shift = h.shift; ix->max_gval = h.max_gval; ix->sort_passes = h.sort_passes; ix->shifts = h.shifts;
  ix->max_len = h.max_len;
  keep_pool_warm(device);
  void* slot[kImgArrays];
  image_slots(ix, slot);
  const char* img = static_cast<const char*>(d_image);
  int rc = BCU_OK;
  for (int k = 0; k < kImgArrays && rc == BCU_OK; ++k) {
    if (!h.size[k]) continue;
    void* p = nullptr;
    if (cudaMallocAsync(&p, h.size[k], stream) != cudaSuccess) { rc = BCU_E_NOMEM; break; }
    *static_cast<void**>(slot[k]) = p;
    if (cudaMemcpyAsync(p, img + h.offset[k], h.size[k], cudaMemcpyDeviceToDevice, stream) != cudaSuccess) rc = BCU_E_CUDA;
  }
  if (rc == BCU_OK && cudaStreamSynchronize(stream) != cudaSuccess) rc = BCU_E_CUDA;
  if (rc != BCU_OK) {
    set_error("bcu_index_import_dev: %s", rc == BCU_E_NOMEM ? "out of device memory" : cudaGetErrorString(cudaGetLastError()));
    free_index_members(ix);
    delete ix;
    return rc;
  }
  if (ix->n && ix->n_comp == 1) {
    // The bin layout is derived data: rebuilt from the imported arrays rather than shipped. It needs the rows in
    // (group, low) order, which the image only holds when the index has ONE length class; an imported index with
    // several classes answers every batch through the general path.
    std::vector<GroupDesc> table((size_t)ix->n_groups);
    const uint64_t bytes_before = ix->bytes;
    if (cudaMemcpyAsync(table.data(), ix->d_groups, table.size() * sizeof(GroupDesc), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) rc = BCU_E_CUDA;
    std::vector<uint32_t> gval(ix->n_groups), begin(ix->n_groups + 1), cmax(ix->n_groups);
    for (uint32_t g = 0; g < ix->n_groups; ++g) {
      gval[g] = table[g].gval;
      begin[g] = table[g].row_begin;
      cmax[g] = (uint32_t)std::min<uint64_t>(((uint64_t)table[g].nb << table[g].shift) - 1, 0xffffffffull);
    }
    begin[ix->n_groups] = (uint32_t)ix->n;
    if (rc == BCU_OK && cudaMallocAsync((void**)&ix->d_bn_low, (ix->n + 4) * 4, stream) != cudaSuccess) rc = BCU_E_NOMEM;
    if (rc == BCU_OK) {
      split_low_kernel<<<(unsigned)((ix->n + 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(ix->d_lowhigh, ix->n + 4, ix->d_bn_low);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      ix->d_bn_high = ix->d_high;
      ix->d_bn_id = ix->d_id;
      rc = build_bin_layout(ix, gval.data(), begin.data(), cmax.data(), stream);
    }
    ix->bytes = bytes_before;  // `bytes` travels in the header and already counts the exporter's layout
    if (rc != BCU_OK) {
      free_index_members(ix);
      delete ix;
      return rc;
    }
  }
  if (ix->n) {  // derived data as well: the stab lists of the long-range emit
    const uint64_t bytes_before = ix->bytes;
    rc = build_long_lists(ix, stream);
    ix->bytes = bytes_before;
    if (rc != BCU_OK) {
      free_index_members(ix);
      delete ix;
      return rc;
    }
  }
  *out = ix;
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
  info->n_components = ix->n_comp;
  info->bin_shift = ix->shift;
  info->sort_passes = ix->sort_passes;
  info->n_bins = ix->n_bins;
  info->device_bytes = ix->bytes;
  info->device = ix->device;
  info->binned_tiles = (int32_t)ix->bn_bins;
  return BCU_OK;
}
