// Standalone validation of the tcgen05 building blocks (tools/, test infrastructure):
//   D[128 x 128] = A[128 x 128] * W[128 x 128]^T with kind::tf32, (a) one pass, (b) 3xTF32 split,
//   operands staged by threads in the k-unit-major canonical layout, W via cp.async.bulk,
//   accumulator read back with tcgen05.ld, plus a TMEM st/ld round trip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tc_test.cu -o tools/tc_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../dig_b200/csrc/tc05.cuh"
using namespace tc05;

constexpr int M = 128, N = 128, K = 64;

// packs W[N][K] (row-major) into [K/4][N][4] hi and lo planes
__global__ void pack_w(const float* w, float* hi, float* lo) {
  int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N * K) return;
  int n = id / K, k = id % K;
  float h, l;
  split_tf32(w[id], h, l);
  size_t o = ((size_t)(k / 4) * N + n) * 4 + (k % 4);
  hi[o] = h; lo[o] = l;
}

struct Smem {
  float a_hi[K / 4 * M * 4];
  float a_lo[K / 4 * M * 4];
  float w_hi[K / 4 * N * 4];
  float w_lo[K / 4 * N * 4];
  uint64_t bar_w, bar_d;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1) tc_kernel(const float* a, const float* w_hi, const float* w_lo, float* d1,
                                                   float* d3, float* stash_out, float* d4) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&s.bar_w, 1); mbar_init(&s.bar_d, 1); mbar_fence_init(); }
  if (warp == 4) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_base;
  // stage A (threads 0..127 own a row each): split + canonical layout
  if (tid < 128) {
    for (int k = 0; k < K; k += 4) {
      float4 v = *reinterpret_cast<const float4*>(a + (size_t)tid * K + k);
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      *reinterpret_cast<float4*>(s.a_hi + ((k / 4) * M + tid) * 4) = h;
      *reinterpret_cast<float4*>(s.a_lo + ((k / 4) * M + tid) * 4) = l;
    }
    fence_async_smem();
  }
  if (tid == 128) {  // producer: bulk copies of the packed weights
    mbar_arrive_expect_tx(&s.bar_w, 2 * N * K * 4);
    bulk_g2s(s.w_hi, w_hi, N * K * 4, &s.bar_w);
    bulk_g2s(s.w_lo, w_lo, N * K * 4, &s.bar_w);
  }
  __syncthreads();
  const uint32_t idesc = idesc_tf32(M, N);
  if (tid == 160) {  // MMA issuer
    mbar_wait(&s.bar_w, 0);
    tc_fence_after();
    // pass (a): single TF32 product into columns [0,128)
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da = smem_desc(smem_u32(s.a_hi) + ks * 2 * M * 16, M * 16, 128);
      uint64_t db = smem_desc(smem_u32(s.w_hi) + ks * 2 * N * 16, N * 16, 128);
      mma_tf32(tm, da, db, idesc, ks > 0);
    }
    // pass (b): 3xTF32 into columns [128,256): lo*hi + hi*lo + hi*hi
    for (int term = 0; term < 3; ++term) {
      const float* pa = term == 0 ? s.a_lo : s.a_hi;
      const float* pb = term == 1 ? s.w_lo : s.w_hi;
      for (int ks = 0; ks < K / 8; ++ks) {
        uint64_t da = smem_desc(smem_u32(pa) + ks * 2 * M * 16, M * 16, 128);
        uint64_t db = smem_desc(smem_u32(pb) + ks * 2 * N * 16, N * 16, 128);
        mma_tf32(tm + 128, da, db, idesc, (term | ks) != 0);
      }
    }
    mma_commit(&s.bar_d);
  }
  if (tid < 128) {   // A operand copies in TMEM: hi at columns [256, 256+K), lo at [384, 384+K)
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    for (int k = 0; k < K; k += 16) {
      uint32_t h[16], l[16];
      for (int i = 0; i < 16; ++i) {
        float hh, ll;
        split_tf32(a[(size_t)tid * K + k + i], hh, ll);
        h[i] = __float_as_uint(hh); l[i] = __float_as_uint(ll);
      }
      tmem_st16(lane_addr + 256 + k, h);
      tmem_st16(lane_addr + 384 + k, l);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  if (tid < 128) {
    mbar_wait(&s.bar_d, 0);
    tc_fence_after();
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 256; c += 16) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c, r);
      tmem_ld_wait();
      float* dst = (c < 128 ? d1 : d3) + (size_t)tid * N + (c & 127);
      for (int i = 0; i < 16; ++i) dst[i] = __uint_as_float(r[i]);
    }
    // TMEM stash round trip: write tid*1000 + col into columns [0,16) and read back
    uint32_t v[16], q[16];
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint((float)(tid * 1000 + i));
    tmem_st16(lane_addr, v);
    tmem_st_wait();
    tmem_ld16(lane_addr, q);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) stash_out[tid * 16 + i] = __uint_as_float(q[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // second round: 4-term product with the A operand read from TMEM (TS form) into columns [0,128)
  if (tid == 160) {
    for (int term = 0; term < 4; ++term) {
      const uint32_t ta = tm + (term == 0 || term == 1 ? 384u : 256u);      // lo, lo, hi, hi
      const float* pb = (term == 0 || term == 2) ? s.w_lo : s.w_hi;          // lo*lo, lo*hi, hi*lo, hi*hi
      for (int ks = 0; ks < K / 8; ++ks) {
        uint64_t db = smem_desc(smem_u32(pb) + ks * 2 * N * 16, N * 16, 128);
        mma_tf32_ts(tm, ta + ks * 8, db, idesc, (term | ks) != 0);
      }
    }
    mma_commit(&s.bar_d);
  }
  if (tid < 128) {
    mbar_wait(&s.bar_d, 1);
    tc_fence_after();
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 128; c += 16) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c, r);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) d4[(size_t)tid * N + c + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int main() {
  std::vector<float> a(M * K), w(N * K);
  srand(1);
  for (auto& x : a) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : w) x = (float)rand() / RAND_MAX * 2 - 1;
  float *da, *dw, *dhi, *dlo, *d1, *d3, *st, *d4;
  cudaMalloc(&da, M * K * 4); cudaMalloc(&dw, N * K * 4); cudaMalloc(&dhi, N * K * 4); cudaMalloc(&dlo, N * K * 4);
  cudaMalloc(&d1, M * N * 4); cudaMalloc(&d3, M * N * 4); cudaMalloc(&st, 128 * 16 * 4); cudaMalloc(&d4, M * N * 4);
  cudaMemcpy(da, a.data(), M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), N * K * 4, cudaMemcpyHostToDevice);
  pack_w<<<(N * K + 255) / 256, 256>>>(dw, dhi, dlo);
  cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  tc_kernel<<<1, 192, sizeof(Smem)>>>(da, dhi, dlo, d1, d3, st, d4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  unsigned int to = 0;
  cudaMemcpyFromSymbol(&to, g_mbar_timeout, sizeof(to));
  printf("mbarrier timeouts: %u\n", to);
  std::vector<float> h1(M * N), h3(M * N), hs(128 * 16), h4(M * N);
  cudaMemcpy(h4.data(), d4, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h1.data(), d1, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3.data(), d3, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs.data(), st, 128 * 16 * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e3 = 0, e4 = 0, ef = 0, ref_max = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)a[i * K + k] * (double)w[j * K + k];
      ref_max = fmax(ref_max, fabs(r));
      e1 = fmax(e1, fabs(h1[i * N + j] - r));
      e3 = fmax(e3, fabs(h3[i * N + j] - r));
      e4 = fmax(e4, fabs(h4[i * N + j] - r));
      { float f = 0; for (int k = 0; k < K; ++k) f = fmaf(a[i * K + k], w[j * K + k], f); ef = fmax(ef, fabs(f - r)); }
    }
  int bad = 0;
  for (int t = 0; t < 128; ++t) for (int i = 0; i < 16; ++i) bad += hs[t * 16 + i] != (float)(t * 1000 + i);
  printf("ref max %.4f | 1xTF32 max abs err %.3e (rel %.3e) | 3xTF32 max abs err %.3e (rel %.3e) | stash mismatches %d\n",
         ref_max, e1, e1 / ref_max, e3, e3 / ref_max, bad);
  printf("4-term TS (A in TMEM) max abs err %.3e (rel %.3e) | plain fp32 fma chain rel %.3e\n", e4, e4 / ref_max, ef / ref_max);
  printf("sample d3[0][0..3] = %f %f %f %f\n", h3[0], h3[1], h3[2], h3[3]);
  return (e == cudaSuccess && to == 0 && e3 / ref_max < 5e-6 && bad == 0) ? 0 : 1;
}
