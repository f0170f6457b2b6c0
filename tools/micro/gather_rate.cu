// Microbenchmark: how many random 8/16-byte reads per second can a B200 serve from an L2-resident table?
// This is the "roofline" that actually bounds the sparse probe kernel (2.3 uncoalesced loads per query),
// as opposed to the HBM stream roofline. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}
template <int BYTES, int DEP>
__global__ void gather(const uint4* __restrict__ table, uint64_t entries, uint64_t n, unsigned long long* out) {
  uint64_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t idx = mix(i) % entries;
#pragma unroll
    for (int d = 0; d < DEP; ++d) {
      if (BYTES == 8) {
        uint2 v = reinterpret_cast<const uint2*>(table)[idx];
        acc += v.x; idx = (mix(i + d + 1) + v.y) % entries;
      } else if (BYTES == 16) {
        uint4 v = table[idx];
        acc += v.x + v.z; idx = (mix(i + d + 1) + v.y) % entries;
      } else {  // 32 bytes: one 256-bit load (LDG.E.256, sm_100+)
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                     : "l"(reinterpret_cast<const char*>(table) + idx * 32));
        acc += r0 + r2 + r4 + r6 + r7; idx = (mix(i + d + 1) + r1 + r3 + r5) % entries;
      }
    }
  }
  if (acc == 0x1234567) *out = acc;
}
template <int BYTES, int DEP> void run(const uint4* t, uint64_t bytes, uint64_t n, unsigned long long* out, const char* label) {
  uint64_t entries = bytes / BYTES;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    gather<BYTES, DEP><<<148 * 8, 256>>>(t, entries, n, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-34s table %4llu MB: %6.1f us for %llu x %d loads -> %.1f G loads/s\n", label,
         (unsigned long long)(bytes >> 20), ms * 1e3, (unsigned long long)n, DEP, n * (double)DEP / ms / 1e6);
}
int main() {
  uint4* t; unsigned long long* out;
  cudaMalloc(&t, 1ull << 30); cudaMemset(t, 1, 1ull << 30); cudaMalloc(&out, 8);
  const uint64_t n = 20000000;
  for (uint64_t mb : {8ull, 24ull, 48ull, 96ull, 400ull}) {
    run<8, 1>(t, mb << 20, n, out, "8 B random, independent");
    run<16, 1>(t, mb << 20, n, out, "16 B random, independent");
    run<8, 2>(t, mb << 20, n, out, "8 B random, chain of 2 dependent");
    run<32, 1>(t, mb << 20, n, out, "32 B random (256-bit), independent");
  }
  return 0;
}
