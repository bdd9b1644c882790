// Issue-rate microbenchmark of tcgen05.mma (test infrastructure): one CTA per SM issues REPS back-to-back MMAs on
// fixed shared-memory operands (contents irrelevant) and reports cycles per instruction for several operand layouts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mma_rate.cu -o tools/mma_rate
#include <cstdio>
#include <cstdlib>
#include "../dig_b200/csrc/tc05.cuh"
using namespace tc05;

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = smem_desc(saddr, lbo, sbo);
  d |= (uint64_t)layout << 61;
  return d;
}

// mode 0: f16 M128 N128 K16, k-unit stride 129 (padded, ours)   1: same, stride 128
// mode 2: f16 N=256 (B = 256 rows)                               3: tf32 M128 N128 K8 stride 129
// mode 4: f16 N128, SWIZZLE_128B K-major (64-element rows of 128 B, SBO = 1024)
// mode 5: f16 N=64
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(&tmem_base, 512);
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0 halves
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t a = smem_u32(smem), b = a + 80 * 1024;
    const int n = mode == 2 ? 256 : (mode == 5 ? 64 : 128);
    const uint32_t idesc = mode == 3 ? idesc_tf32(128, n) : idesc_f16(128, n);
    const uint32_t aku = (mode == 1) ? 128 : 129;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int ks = r & 7;     // walk over 8 k-steps like a K = 128 layer
      uint64_t da, db;
      if (mode == 4) {
        da = desc_sw(a + ks * 32, 16, 1024, 2);
        db = desc_sw(b + ks * 32, 16, 1024, 2);
      } else {
        da = smem_desc(a + ks * 2 * aku * 16, aku * 16, 128);
        db = smem_desc(b + ks * 2 * n * 16, n * 16, 128);
      }
      if (mode == 3) mma_tf32(tmem_base, da, db, idesc, r != 0);
      else mma_f16(tmem_base, da, db, idesc, r != 0);
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"f16 N128 K16 stride129 (chain layout)", "f16 N128 K16 stride128", "f16 N256 K16",
                         "tf32 N128 K8 stride129", "f16 N128 K16 SWIZZLE_128B", "f16 N64 K16"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int grid : {1, 148}) {
      const int reps = 2048;
      rate_kernel<<<grid, 128, 200 * 1024>>>(mode, reps, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0};
      cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
      printf("%-40s grid %3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA  (%s)\n", names[mode], grid,
             (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(e));
    }
  }
  return 0;
}
