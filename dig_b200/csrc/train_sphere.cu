// Fused backward kernels of the SphereNet / DimeNet++ training path (sm_100a).
//
//   sphere_triplet_gather_bwd : backward of   m[e] = sum_{t in trip(e)} x_down[kj(t)] * lin_sbf2(sbf_p[t]) * lin_t2(t_p[t])
//                               (reference spherenet.py:163-171; forward = sphere_triplet_gather_kernel, spherenet_tc.cu).
//                               One pass over the triplets produces d x_down (atomics: kj is not sorted), d sbf_p, d t_p and
//                               the two [64, 8] weight gradients, instead of ~12 elementwise / GEMM launches over [T, 64].
#include "common.cuh"

namespace dig3d {

// Sum 16 per-lane values over the warp with 16 shuffles (recursive halving): afterwards lane L holds the total of
// value L >> 1 (both lanes of a pair hold the same number).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// One warp per (j -> i) edge e, lanes = channels (lane, lane + 32) as in the forward kernel; persistent CTAs so that
// the weight gradients are accumulated in registers over many edges and flushed once per CTA.
template <bool TORSION>
__global__ void __launch_bounds__(256, 2)
sphere_triplet_gather_bwd_kernel(const float* __restrict__ dm, const float* __restrict__ x_down,
                                 const float* __restrict__ sbf_p, const float* __restrict__ t_p,
                                 const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                 const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ trip_ptr, int n_edges,
                                 const float* __restrict__ w_sbf2, const float* __restrict__ w_t2,
                                 float* __restrict__ dx_down, float* __restrict__ d_sbf_p, float* __restrict__ d_t_p,
                                 float* __restrict__ dw_sbf2, float* __restrict__ dw_t2) {
  __shared__ __align__(16) float stage[8][2][64];
  __shared__ float sdw[2][64 * 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int id = threadIdx.x; id < 2 * 64 * 8; id += 256) (&sdw[0][0])[id] = 0.f;
  float ws2[2][8], wt2[2][8], gs2[2][8], gt2[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      ws2[h][q] = __ldg(w_sbf2 + (lane + 32 * h) * 8 + q);
      wt2[h][q] = TORSION ? __ldg(w_t2 + (lane + 32 * h) * 8 + q) : 0.f;
      gs2[h][q] = 0.f;
      gt2[h][q] = 0.f;
    }
  __syncthreads();
  const int n_warps = gridDim.x * 8;
  for (int e = blockIdx.x * 8 + w; e < n_edges; e += n_warps) {
    const int j = src[e], i = dst[e];
    const int base = row_ptr[j], d = row_ptr[j + 1] - base;
    int p_i = d;
    for (int s0 = 0; s0 < d; s0 += 32) {
      const int sl = s0 + lane;
      const unsigned hit = __ballot_sync(0xffffffffu, sl < d && src[base + sl] == i);
      if (hit) p_i = s0 + __ffs(hit) - 1;
    }
    const int t0 = trip_ptr[e], nt = d - (p_i < d ? 1 : 0);
    const float dm0 = __ldg(dm + (size_t)e * 64 + lane), dm1 = __ldg(dm + (size_t)e * 64 + lane + 32);
    for (int r0 = 0; r0 < nt; r0 += 8) {
      const int n8 = min(8, nt - r0), lim = n8 * 8;
      const float* sp = sbf_p + (size_t)(t0 + r0) * 8;
      const float sa = lane < lim ? __ldg(sp + lane) : 0.f, sb = lane + 32 < lim ? __ldg(sp + lane + 32) : 0.f;
      float ta = 0.f, tb = 0.f;
      if (TORSION) {
        const float* tp = t_p + (size_t)(t0 + r0) * 8;
        ta = lane < lim ? __ldg(tp + lane) : 0.f;
        tb = lane + 32 < lim ? __ldg(tp + lane + 32) : 0.f;
      }
      __syncwarp();
      stage[w][0][lane] = sa; stage[w][0][lane + 32] = sb;
      if (TORSION) { stage[w][1][lane] = ta; stage[w][1][lane + 32] = tb; }
      __syncwarp();
      for (int u = 0; u < n8; ++u) {
        const int r = r0 + u;
        const int kj = base + r + (r >= p_i ? 1 : 0);
        const float x0 = __ldg(x_down + (size_t)kj * 64 + lane), x1 = __ldg(x_down + (size_t)kj * 64 + lane + 32);
        const float4 s0 = *reinterpret_cast<const float4*>(&stage[w][0][u * 8]);
        const float4 s1 = *reinterpret_cast<const float4*>(&stage[w][0][u * 8 + 4]);
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        float tv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float g0 = 0.f, g1 = 0.f, h0 = 1.f, h1 = 1.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { g0 = fmaf(ws2[0][q], sv[q], g0); g1 = fmaf(ws2[1][q], sv[q], g1); }
        if (TORSION) {
          const float4 q0 = *reinterpret_cast<const float4*>(&stage[w][1][u * 8]);
          const float4 q1 = *reinterpret_cast<const float4*>(&stage[w][1][u * 8 + 4]);
          tv[0] = q0.x; tv[1] = q0.y; tv[2] = q0.z; tv[3] = q0.w; tv[4] = q1.x; tv[5] = q1.y; tv[6] = q1.z; tv[7] = q1.w;
          h0 = 0.f; h1 = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) { h0 = fmaf(wt2[0][q], tv[q], h0); h1 = fmaf(wt2[1][q], tv[q], h1); }
        }
        // d x_down[kj] += dm * s * t
        atomicAdd(dx_down + (size_t)kj * 64 + lane, dm0 * g0 * h0);
        atomicAdd(dx_down + (size_t)kj * 64 + lane + 32, dm1 * g1 * h1);
        const float p0 = dm0 * x0, p1 = dm1 * x1;
        const float a0 = p0 * h0, a1 = p1 * h1;      // d(s) per channel
        const float b0 = p0 * g0, b1 = p1 * g1;      // d(t) per channel
        float v[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          gs2[0][q] = fmaf(a0, sv[q], gs2[0][q]);
          gs2[1][q] = fmaf(a1, sv[q], gs2[1][q]);
          v[q] = fmaf(a0, ws2[0][q], a1 * ws2[1][q]);
          if (TORSION) {
            gt2[0][q] = fmaf(b0, tv[q], gt2[0][q]);
            gt2[1][q] = fmaf(b1, tv[q], gt2[1][q]);
            v[8 + q] = fmaf(b0, wt2[0][q], b1 * wt2[1][q]);
          } else {
            v[8 + q] = 0.f;
          }
        }
        const float tot = warp_reduce16(v, lane);
        const int idx = lane >> 1;
        if ((lane & 1) == 0) {
          if (idx < 8) d_sbf_p[(size_t)(t0 + r) * 8 + idx] = tot;
          else if (TORSION) d_t_p[(size_t)(t0 + r) * 8 + idx - 8] = tot;
        }
      }
    }
  }
  // CTA reduction of the weight gradients, then one atomic flush
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      atomicAdd(&sdw[0][(lane + 32 * h) * 8 + q], gs2[h][q]);
      if (TORSION) atomicAdd(&sdw[1][(lane + 32 * h) * 8 + q], gt2[h][q]);
    }
  __syncthreads();
  for (int id = threadIdx.x; id < 64 * 8; id += 256) {
    atomicAdd(dw_sbf2 + id, sdw[0][id]);
    if (TORSION) atomicAdd(dw_t2 + id, sdw[1][id]);
  }
}

}  // namespace dig3d

using namespace dig3d;

extern "C" int dig3d_sphere_triplet_gather_bwd(const float* dm, const float* x_down, const float* sbf_p, const float* t_p,
                                               const int32_t* src, const int32_t* dst, const int32_t* row_ptr,
                                               const int32_t* trip_ptr, int64_t n_edges, const float* w_sbf2,
                                               const float* w_t2, float* dx_down, float* d_sbf_p, float* d_t_p,
                                               float* dw_sbf2, float* dw_t2, void* stream) {
  DIG3D_REQUIRE(dm && x_down && sbf_p && src && dst && row_ptr && trip_ptr && w_sbf2 && dx_down && d_sbf_p && dw_sbf2,
                "sphere_triplet_gather_bwd: null pointer");
  DIG3D_REQUIRE((t_p != nullptr) == (w_t2 != nullptr) && (t_p != nullptr) == (d_t_p != nullptr) &&
                    (t_p != nullptr) == (dw_t2 != nullptr),
                "sphere_triplet_gather_bwd: the torsion arguments must be all set or all null");
  if (n_edges == 0) return DIG3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ceil_div(n_edges, 8);
  if (grid > 296) grid = 296;
  if (t_p)
    sphere_triplet_gather_bwd_kernel<true><<<grid, 256, 0, st>>>(dm, x_down, sbf_p, t_p, src, dst, row_ptr, trip_ptr,
                                                                 (int)n_edges, w_sbf2, w_t2, dx_down, d_sbf_p, d_t_p,
                                                                 dw_sbf2, dw_t2);
  else
    sphere_triplet_gather_bwd_kernel<false><<<grid, 256, 0, st>>>(dm, x_down, sbf_p, t_p, src, dst, row_ptr, trip_ptr,
                                                                  (int)n_edges, w_sbf2, w_t2, dx_down, d_sbf_p, d_t_p,
                                                                  dw_sbf2, dw_t2);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}
