// FP32 issue-rate microbenchmark (test infrastructure): per-SM throughput of scalar FFMA (three register sources) against
// the packed fma.rn.f32x2 (FFMA2) of sm_100, at several warps per scheduler.  Each thread runs ILP independent
// accumulator chains so that latency is hidden; the result is FMA lanes per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ffma_rate.cu -o tools/ffma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

constexpr int ILP = 8;

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  float2 v = make_float2(lo, hi);
  return *reinterpret_cast<unsigned long long*>(&v);
}

// mode 0: scalar FFMA, acc = a * b + acc with a, b, acc all registers (distinct per chain)
// mode 1: FFMA2, same shape on register pairs
template <int MODE>
__global__ void rate_kernel(int reps, const float* __restrict__ in, float* __restrict__ out, long long* cyc) {
  float a[ILP], b[ILP], acc[2 * ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = in[i] + threadIdx.x; b[i] = in[ILP + i] - 1e-6f * threadIdx.x; }
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) acc[i] = in[i & 7];
  // operand pairs formed ONCE (64-bit registers), so the timed loop holds nothing but the FMAs
  unsigned long long a2[ILP / 2], b2[ILP / 2], c2[ILP];
#pragma unroll
  for (int i = 0; i < ILP / 2; ++i) {
    a2[i] = pack2(a[2 * i], a[2 * i + 1]);
    b2[i] = pack2(b[2 * i], b[2 * i + 1]);
  }
#pragma unroll
  for (int i = 0; i < ILP; ++i) c2[i] = pack2(acc[2 * i], acc[2 * i + 1]);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 2 * ILP; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a[i & 7]), "f"(b[(i + 3) & 7]));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c2[i]) : "l"(a2[i & 3]), "l"(b2[(i + 1) & 3]));
    } else {     // MODE 2: scalar FFMA with one operand shared by all chains of the iteration (reuse-cache friendly, GEMM-like)
#pragma unroll
      for (int i = 0; i < 2 * ILP; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a[r & 7 ? 0 : 1]), "f"(b[i & 7]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float2 v = *reinterpret_cast<float2*>(&c2[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float h_in[16];
  for (int i = 0; i < 16; ++i) h_in[i] = 1.0f + 1e-3f * i;
  float *d_in, *d_out;
  long long* d_cyc;
  cudaMalloc(&d_in, sizeof(h_in));
  cudaMalloc(&d_out, 148 * 1024 * 4);
  cudaMalloc(&d_cyc, 8);
  cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
  const int reps = 4096;
  const char* names[] = {"FFMA", "FFMA2", "FFMA-r"};
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      if (mode == 0) rate_kernel<0><<<148, threads>>>(reps, d_in, d_out, d_cyc);
      else if (mode == 1) rate_kernel<1><<<148, threads>>>(reps, d_in, d_out, d_cyc);
      else rate_kernel<2><<<148, threads>>>(reps, d_in, d_out, d_cyc);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
      const double fmas = (double)reps * 2 * ILP * threads;   // FMA lanes per CTA (= per SM)
      printf("%-6s %4d threads/SM: %.1f FMA lanes / clk / SM  (%lld cycles, %s)\n", names[mode], threads,
             fmas / (double)c, c, cudaGetErrorString(e));
    }
  return 0;
}
