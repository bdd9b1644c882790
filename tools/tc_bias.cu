// Is the fp32 accumulation inside tcgen05.mma biased (round-toward-zero)?  Measures the mean SIGNED relative
// error of a 3xTF32 product  D = A W^T  (M=N=128, K=128 via two K=64 halves), for:
//   (a) one TMEM accumulator for everything,
//   (b) separate accumulators: hi*hi per K-quarter (4) + one for the two correction terms, summed in fp32 (RN) afterwards.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../dig_b200/csrc/tc05.cuh"
using namespace tc05;
constexpr int M = 128, N = 128, K = 64;
__global__ void pack_w(const float* w, float* hi, float* lo) {
  int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N * K) return;
  int n = id / K, k = id % K;
  float h, l; split_tf32(w[id], h, l);
  size_t o = ((size_t)(k / 4) * N + n) * 4 + (k % 4);
  hi[o] = h; lo[o] = l;
}
struct Smem { float a_hi[K / 4 * M * 4], a_lo[K / 4 * M * 4], w_hi[K / 4 * N * 4], w_lo[K / 4 * N * 4]; uint64_t bar_w, bar_d; uint32_t tmem_base; };
__global__ void __launch_bounds__(192, 1) k(const float* a, const float* w_hi, const float* w_lo, float* d_one, float* d_split) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&s.bar_w, 1); mbar_init(&s.bar_d, 1); mbar_fence_init(); }
  if (warp == 4) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = s.tmem_base;
  if (tid < 128) {
    for (int kk = 0; kk < K; kk += 4) {
      float4 v = *reinterpret_cast<const float4*>(a + (size_t)tid * K + kk), h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      *reinterpret_cast<float4*>(s.a_hi + ((kk / 4) * M + tid) * 4) = h;
      *reinterpret_cast<float4*>(s.a_lo + ((kk / 4) * M + tid) * 4) = l;
    }
    fence_async_smem();
  }
  if (tid == 128) { mbar_arrive_expect_tx(&s.bar_w, 2 * N * K * 4); bulk_g2s(s.w_hi, w_hi, N * K * 4, &s.bar_w); bulk_g2s(s.w_lo, w_lo, N * K * 4, &s.bar_w); }
  __syncthreads();
  const uint32_t idesc = idesc_tf32(M, N);
  if (tid == 160) {
    mbar_wait(&s.bar_w, 0); tc_fence_after();
    auto DA = [&](const float* p, int ks) { return smem_desc(smem_u32(p) + ks * 2 * M * 16, M * 16, 128); };
    auto DB = [&](const float* p, int ks) { return smem_desc(smem_u32(p) + ks * 2 * N * 16, N * 16, 128); };
    // (a) single accumulator, columns [0,128)
    for (int ks = 0; ks < K / 8; ++ks) {
      mma_tf32(tm, DA(s.a_lo, ks), DB(s.w_hi, ks), idesc, ks > 0);
      mma_tf32(tm, DA(s.a_hi, ks), DB(s.w_lo, ks), idesc, 1);
      mma_tf32(tm, DA(s.a_hi, ks), DB(s.w_hi, ks), idesc, 1);
    }
    // (b) hi*hi in two accumulators (K halves) [128,256) [256,384); corrections in [384,512)
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t dm = tm + 128 + (ks / (K / 16)) * 128;
      mma_tf32(dm, DA(s.a_hi, ks), DB(s.w_hi, ks), idesc, (ks % (K / 16)) != 0);
      mma_tf32(tm + 384, DA(s.a_lo, ks), DB(s.w_hi, ks), idesc, ks > 0);
      mma_tf32(tm + 384, DA(s.a_hi, ks), DB(s.w_lo, ks), idesc, 1);
    }
    mma_commit(&s.bar_d);
  }
  if (tid < 128) {
    mbar_wait(&s.bar_d, 0); tc_fence_after();
    const uint32_t la = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 128; c += 16) {
      uint32_t r[16], p0[16], p1[16], pc[16];
      tmem_ld16(la + c, r); tmem_ld16(la + 128 + c, p0); tmem_ld16(la + 256 + c, p1); tmem_ld16(la + 384 + c, pc);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) {
        d_one[(size_t)tid * N + c + i] = __uint_as_float(r[i]);
        d_split[(size_t)tid * N + c + i] = __fadd_rn(__fadd_rn(__uint_as_float(p0[i]), __uint_as_float(p1[i])), __uint_as_float(pc[i]));
      }
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}
int main() {
  std::vector<float> a(M * K), w(N * K);
  srand(3);
  for (auto& x : a) x = (float)rand() / RAND_MAX;          // all positive: every partial sum grows, bias is visible
  for (auto& x : w) x = (float)rand() / RAND_MAX;
  float *da, *dw, *dhi, *dlo, *d1, *d2;
  cudaMalloc(&da, M * K * 4); cudaMalloc(&dw, N * K * 4); cudaMalloc(&dhi, N * K * 4); cudaMalloc(&dlo, N * K * 4);
  cudaMalloc(&d1, M * N * 4); cudaMalloc(&d2, M * N * 4);
  cudaMemcpy(da, a.data(), M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), N * K * 4, cudaMemcpyHostToDevice);
  pack_w<<<(N * K + 255) / 256, 256>>>(dw, dhi, dlo);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  k<<<1, 192, sizeof(Smem)>>>(da, dhi, dlo, d1, d2);
  printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<float> h1(M * N), h2(M * N);
  cudaMemcpy(h1.data(), d1, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), d2, M * N * 4, cudaMemcpyDeviceToHost);
  double b1 = 0, b2 = 0, bf = 0, m1 = 0, m2 = 0, mf = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    double r = 0; float f = 0;
    for (int kk = 0; kk < K; ++kk) { r += (double)a[i * K + kk] * w[j * K + kk]; f = fmaf(a[i * K + kk], w[j * K + kk], f); }
    double e1 = (h1[i * N + j] - r) / r, e2 = (h2[i * N + j] - r) / r, ef = (f - r) / r;
    b1 += e1; b2 += e2; bf += ef; m1 = fmax(m1, fabs(e1)); m2 = fmax(m2, fabs(e2)); mf = fmax(mf, fabs(ef));
  }
  const double n = M * N;
  printf("positive data, K=%d: mean signed rel err | one accumulator %.3e (max %.3e) | split accumulators %.3e (max %.3e) | fp32 fma chain %.3e (max %.3e)\n",
         K, b1 / n, m1, b2 / n, m2, bf / n, mf);
  return 0;
}
