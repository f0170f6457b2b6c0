// Device-wide single-pass scans built on lookback.cuh:
//   exclusive_sum_u32      -- digit-histogram prefix sums for the radix sort (K1)
//   segmented_running_max  -- the "running max-end" array of the flat index (K2): for each sorted row
//                             the max of `high` over the rows of the same segment (component, group) up to
//                             and including it
// HBM roofline: 1 read + 1 write of the array (4 B + 4 B per element).
#include "common.cuh"
#include "lookback.cuh"

namespace bcu {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;  // one 128-bit load per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
    exclusive_sum_u32_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n,
                             uint64_t* status, uint32_t* ticket) {
  __shared__ uint64_t s_scan[kScanThreads / 32 + 1];
  __shared__ uint32_t s_tile;
  __shared__ uint64_t s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  if (base + kScanItems <= n) {
    uint4 t = *reinterpret_cast<const uint4*>(in + base);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n) ? in[base + k] : 0u;
  }
  uint64_t thread_sum = (uint64_t)v[0] + v[1] + v[2] + v[3];
  uint64_t block_total;
  uint64_t excl = block_exclusive_scan<SumOp, kScanThreads>(thread_sum, s_scan, &block_total);
  if (threadIdx.x < 32) {
    uint64_t e = lookback_exclusive<SumOp>(status, tile, block_total);
    if (threadIdx.x == 0) s_excl = e;
  }
  __syncthreads();
  uint32_t run = (uint32_t)(s_excl + excl);
  uint32_t o[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { o[k] = run; run += v[k]; }
  if (base + kScanItems <= n) {
    *reinterpret_cast<uint4*>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) out[base + k] = o[k];
  }
}

__global__ void __launch_bounds__(kScanThreads)
    segmented_running_max_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ val,
                                 uint32_t* __restrict__ out, uint64_t n, uint64_t* status,
                                 uint32_t* ticket) {
  __shared__ uint64_t s_scan[kScanThreads / 32 + 1];
  __shared__ uint32_t s_tile;
  __shared__ uint64_t s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint64_t e[kScanItems];  // SegMaxOp elements: head flag in bit 32, value in the low word
  uint64_t prev_seg = 0;
  if (base > 0 && base < n) prev_seg = keys[base - 1];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) {
      uint64_t g = keys[base + k];  // segment key: rows of one segment carry the same value
      bool head = (base + k == 0) || (g != prev_seg);
      e[k] = ((uint64_t)head << 32) | val[base + k];
      prev_seg = g;
    } else {
      e[k] = SegMaxOp::identity();
    }
  }
  uint64_t incl[kScanItems];
  incl[0] = e[0];
#pragma unroll
  for (int k = 1; k < kScanItems; ++k) incl[k] = SegMaxOp::combine(incl[k - 1], e[k]);
  uint64_t block_total;
  uint64_t excl = block_exclusive_scan<SegMaxOp, kScanThreads>(incl[kScanItems - 1], s_scan, &block_total);
  if (threadIdx.x < 32) {
    uint64_t p = lookback_exclusive<SegMaxOp>(status, tile, block_total);
    if (threadIdx.x == 0) s_excl = p;
  }
  __syncthreads();
  uint64_t carry = SegMaxOp::combine(s_excl, excl);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) out[base + k] = (uint32_t)SegMaxOp::combine(carry, incl[k]);
}

static int alloc_status(uint64_t tiles, cudaStream_t stream, uint64_t** status, uint32_t** ticket) {
  void* p = nullptr;
  size_t bytes = (tiles + 1) * sizeof(uint64_t);
  BCU_CUDA(cudaMallocAsync(&p, bytes, stream));
  BCU_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
  *status = reinterpret_cast<uint64_t*>(p);
  *ticket = reinterpret_cast<uint32_t*>(*status + tiles);
  return BCU_OK;
}

int exclusive_sum_u32(const uint32_t* d_in, uint32_t* d_out, uint64_t n, cudaStream_t stream) {
  if (n == 0) return BCU_OK;
  uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  uint64_t* status; uint32_t* ticket;
  BCU_TRY(alloc_status(tiles, stream, &status, &ticket));
  exclusive_sum_u32_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(d_in, d_out, n, status, ticket);
  BCU_LAUNCHED();
  BCU_CUDA(cudaFreeAsync(status, stream));
  return BCU_OK;
}

int segmented_running_max(const uint64_t* d_keys, const uint32_t* d_val, uint32_t* d_out, uint64_t n,
                          cudaStream_t stream) {
  if (n == 0) return BCU_OK;
  uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  uint64_t* status; uint32_t* ticket;
  BCU_TRY(alloc_status(tiles, stream, &status, &ticket));
  segmented_running_max_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(d_keys, d_val, d_out, n,
                                                                             status, ticket);
  BCU_LAUNCHED();
  BCU_CUDA(cudaFreeAsync(status, stream));
  return BCU_OK;
}

}  // namespace bcu
