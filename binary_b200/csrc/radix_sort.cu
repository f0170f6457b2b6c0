// K1 -- stable LSD radix sort of interval records keyed by (group << 32 | low), payload = target id.
//
// Replaces the reference's per-node red-black insertion (rb_tree.hpp:145-149,304-344;
// interval_tree.hpp:230-260) as the ordering step of the build. Only digits in which the keys
// actually differ are run (hg38: 28 coordinate bits + 5 group bits -> 5 passes of <= 8 bits).
// Per pass: tile digit histogram -> device-wide exclusive scan (decoupled look-back, scan.cu) ->
// stable scatter using warp match-any ranking. HBM roofline per pass: 12 B read (hist reads keys
// only: 8 B) + 12 B read + 12 B write per record.
#include "common.cuh"

namespace bcu {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 records per CTA
constexpr int kRadix = 256;

// hist[digit * n_tiles + tile] = number of keys of `tile` whose digit equals `digit`
__global__ void __launch_bounds__(kSortThreads)
    sort_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t mask,
                     uint32_t* __restrict__ hist, uint32_t n_tiles) {
  __shared__ uint32_t s_hist[kRadix];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    uint64_t i = base + (uint64_t)k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[(uint32_t)(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = s_hist[threadIdx.x];
}

// Stable scatter. Warp w owns the contiguous slice [w*256, (w+1)*256) of the tile and walks it in 8
// steps of 32 consecutive records, so (warp, step, lane) order == input order.
__global__ void __launch_bounds__(kSortThreads)
    sort_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                        uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint64_t n,
                        int shift, uint32_t mask, const uint32_t* __restrict__ scanned,
                        uint32_t n_tiles) {
  __shared__ uint32_t s_cnt[kSortWarps][kRadix];  // per-warp running digit counts -> warp bases
  __shared__ uint32_t s_base[kRadix];             // global destination of the tile's first key per digit
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) s_cnt[w][threadIdx.x] = 0;
  s_base[threadIdx.x] = scanned[(uint64_t)threadIdx.x * n_tiles + blockIdx.x];
  __syncthreads();

  const uint64_t wbase = (uint64_t)blockIdx.x * kSortTile + (uint64_t)warp * (32 * kSortItems);
  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    uint64_t i = wbase + (uint64_t)k * 32 + lane;
    bool live = i < n;
    key[k] = live ? keys_in[i] : ~0ull;
    val[k] = live ? vals_in[i] : 0u;
    uint32_t d = (uint32_t)(key[k] >> shift) & mask;
    // lanes holding the same digit (dead lanes are parked in a digit of their own via the live bit)
    unsigned peers = __match_any_sync(0xffffffffu, live ? d : (kRadix + lane));
    int leader = __ffs(peers) - 1;
    uint32_t before = 0;
    if (live && lane == leader) {
      before = s_cnt[warp][d];
      s_cnt[warp][d] = before + __popc(peers);
    }
    before = __shfl_sync(0xffffffffu, before, leader);
    rank[k] = before + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over the warps, per digit (thread d handles digit d)
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = s_cnt[w][threadIdx.x];
      s_cnt[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    uint64_t i = wbase + (uint64_t)k * 32 + lane;
    if (i < n) {
      uint32_t d = (uint32_t)(key[k] >> shift) & mask;
      uint32_t dst = s_base[d] + s_cnt[warp][d] + rank[k];
      keys_out[dst] = key[k];
      vals_out[dst] = val[k];
    }
  }
}

int radix_sort_pairs(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint64_t n,
                     uint64_t varying_bits, cudaStream_t stream, uint64_t** out_keys,
                     uint32_t** out_vals, uint32_t* passes) {
  *out_keys = keys_a;
  *out_vals = vals_a;
  *passes = 0;
  if (n < 2 || varying_bits == 0) return BCU_OK;
  const uint32_t n_tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
  uint32_t *hist = nullptr, *scanned = nullptr;
  const uint64_t hist_len = (uint64_t)kRadix * n_tiles;
  BCU_CUDA(cudaMallocAsync((void**)&hist, hist_len * 4, stream));
  BCU_CUDA(cudaMallocAsync((void**)&scanned, hist_len * 4, stream));
  uint64_t *kin = keys_a, *kout = keys_b;
  uint32_t *vin = vals_a, *vout = vals_b;
  int bit = 0;
  while (bit < 64) {
    if (!((varying_bits >> bit) & 1ull)) { ++bit; continue; }  // skip constant low bits
    // digit = up to 8 bits starting at the first varying bit; stop early at a run of constant bits
    int width = 0;
    while (width < 8 && bit + width < 64) ++width;
    uint64_t window = (varying_bits >> bit) & ((1ull << width) - 1ull);
    if (window == 0) { bit += width; continue; }
    while (width > 1 && !((window >> (width - 1)) & 1ull)) --width;  // trim constant high bits
    const uint32_t mask = (1u << width) - 1u;
    sort_hist_kernel<<<n_tiles, kSortThreads, 0, stream>>>(kin, n, bit, mask, hist, n_tiles);
    BCU_LAUNCHED();
    BCU_TRY(exclusive_sum_u32(hist, scanned, hist_len, stream));
    sort_scatter_kernel<<<n_tiles, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, bit, mask, scanned,
                                                             n_tiles);
    BCU_LAUNCHED();
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
    bit += width;
    ++*passes;
  }
  BCU_CUDA(cudaFreeAsync(hist, stream));
  BCU_CUDA(cudaFreeAsync(scanned, stream));
  *out_keys = kin;
  *out_vals = vin;
  return BCU_OK;
}

}  // namespace bcu
