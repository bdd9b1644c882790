// SphereNet / DimeNet++ interaction blocks (sm_100a), fp32.
//
//   init.forward        spherenet.py:79-91   dimenetpp.py:71-78
//   update_e.forward    spherenet.py:150-182 dimenetpp.py:133-161
//   update_v.forward    spherenet.py:209-216 dimenetpp.py:188-195
//
// One CTA owns a tile of 64 consecutive edges (edges are sorted by target node) and keeps the
// tile's activations in shared memory across the whole chain of linears of the block; weights
// are streamed from L2.  The edge->node scatter of update_v is fused into the producing kernel
// as a segmented reduction over the (sorted) target index, so e2 is never written to HBM.
// The triplet->edge scatter (spherenet.py:171) is a register accumulation over the contiguous
// triplet range of each edge: no atomics, no [T, 64] intermediates.
#include "dense.cuh"

namespace dig3d {

constexpr int H = 128;     // hidden_channels
constexpr int IE = 64;     // int_emb_size
constexpr int NRAD = 6;    // num_radial
constexpr int BE = 8;      // basis_emb_size
constexpr int OE = 256;    // out_emb_channels
constexpr int TM = 64;     // edges per CTA
constexpr int LDA = H + 4;
constexpr int TMV = 32;    // nodes per CTA in update_v
constexpr int LDV = OE + 4;

struct EdgeSmem {
  float buf0[TM * LDA];
  float buf1[TM * LDA];
  float ws[2 * H * LDW];
  float rbf[TM * 8];
  float r8[TM * 8];
  int src[TM];
  int dst[TM];
};

__device__ __forceinline__ void load_rbf_tile(EdgeSmem& s, const float* __restrict__ rbf0, int e0, int rows) {
  for (int id = threadIdx.x; id < TM * 8; id += DT) {
    const int r = id >> 3, n = id & 7;
    s.rbf[id] = (r < rows && n < NRAD) ? __ldg(rbf0 + (size_t)(e0 + r) * NRAD + n) : 0.f;
  }
}

// y[r][c] = sum_n w[c][n] * rbf[r][n]   (K = num_radial, nn.Linear without bias)
__device__ __forceinline__ float rbf_dot(const float* __restrict__ w, int c, const float* rb) {
  float a = 0.f;
#pragma unroll
  for (int n = 0; n < NRAD; ++n) a = fmaf(__ldg(w + c * NRAD + n), rb[n], a);
  return a;
}

// ------------------------------------------------------------------ init_e
__global__ void __launch_bounds__(DT, 2)
sphere_init_e_kernel(const int64_t* __restrict__ z, const int32_t* __restrict__ src,
                     const int32_t* __restrict__ dst, const float* __restrict__ rbf0, int n_edges,
                     dig3d_init_e_weights W, float* __restrict__ e1, float* __restrict__ v_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EdgeSmem& s = *reinterpret_cast<EdgeSmem*>(smem_raw);
  const int e0 = blockIdx.x * TM, rows = min(TM, n_edges - e0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (int r = threadIdx.x; r < TM; r += DT) {
    s.src[r] = (r < rows) ? src[e0 + r] : -1;
    s.dst[r] = (r < rows) ? dst[e0 + r] : -1;
  }
  load_rbf_tile(s, rbf0, e0, rows);
  __syncthreads();
  float acc[4][8];
  zero_acc(acc);
  // cat([x_i, x_j, rbf0]) @ W_lin^T as three K=128 panels          spherenet.py:88
  for (int seg = 0; seg < 3; ++seg) {
    for (int id = threadIdx.x; id < TM * (H / 4); id += DT) {
      const int r = id / (H / 4), c = (id % (H / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        if (seg < 2) {
          const int node = seg == 0 ? s.dst[r] : s.src[r];
          v = __ldg(reinterpret_cast<const float4*>(W.emb + (size_t)z[node] * H + c));
        } else {  // rbf0 = act(lin_rbf_0(rbf))                      spherenet.py:87
          const float* rb = s.rbf + r * 8;
          v.x = swish(rbf_dot(W.w_rbf0, c + 0, rb) + __ldg(W.b_rbf0 + c + 0));
          v.y = swish(rbf_dot(W.w_rbf0, c + 1, rb) + __ldg(W.b_rbf0 + c + 1));
          v.z = swish(rbf_dot(W.w_rbf0, c + 2, rb) + __ldg(W.b_rbf0 + c + 2));
          v.w = swish(rbf_dot(W.w_rbf0, c + 3, rb) + __ldg(W.b_rbf0 + c + 3));
        }
      }
      *reinterpret_cast<float4*>(s.buf0 + r * LDA + c) = v;
    }
    __syncthreads();
    gemm_tile<TM, H, H>(s.buf0, LDA, W.w_lin + seg * H, 3 * H, s.ws, acc);
  }
  // e1 = act(.), e2 = lin_rbf_1(rbf) * e1                            spherenet.py:88-89
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = ty * 4 + p;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tx + 16 * q;
      const float v = swish(acc[p][q] + __ldg(W.b_lin + c));
      s.buf0[r * LDA + c] = v;
      s.buf1[r * LDA + c] = __fmul_rn(rbf_dot(W.w_rbf1, c, s.rbf + r * 8), v);
    }
  }
  __syncthreads();
  tile_store<H>(e1 + (size_t)e0 * H, H, s.buf0, LDA, rows);
  tile_segment_accumulate(s.buf1, LDA, s.dst, rows, v_in, H);
}

// ------------------------------------------------------------------ update_e, part A
__global__ void __launch_bounds__(DT, 2)
sphere_update_e_a_kernel(const float* __restrict__ e1, const float* __restrict__ rbf0, int n_edges,
                         dig3d_update_e_weights W, float* __restrict__ x_ji, float* __restrict__ x_down) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EdgeSmem& s = *reinterpret_cast<EdgeSmem*>(smem_raw);
  const int e0 = blockIdx.x * TM, rows = min(TM, n_edges - e0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  tile_load<H>(s.buf0, LDA, e1 + (size_t)e0 * H, H, rows);
  for (int id = threadIdx.x; id < (TM - rows) * H; id += DT)
    s.buf0[(rows + id / H) * LDA + id % H] = 0.f;
  load_rbf_tile(s, rbf0, e0, rows);
  __syncthreads();
  for (int id = threadIdx.x; id < TM * BE; id += DT) {  // rbf = lin_rbf1(rbf0)   spherenet.py:157
    const int r = id >> 3, m = id & 7;
    s.r8[id] = rbf_dot(W.w_rbf1, m, s.rbf + r * 8);
  }
  float acc[4][8];
  // x_ji = act(lin_ji(x1))                                             spherenet.py:154
  zero_acc(acc);
  gemm_tile<TM, H, H>(s.buf0, LDA, W.w_ji, H, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tx + 16 * q;
      s.buf1[(ty * 4 + p) * LDA + c] = swish(acc[p][q] + __ldg(W.b_ji + c));
    }
  __syncthreads();
  tile_store<H>(x_ji + (size_t)e0 * H, H, s.buf1, LDA, rows);
  // x_kj = act(lin_kj(x1)) * lin_rbf2(rbf)                             spherenet.py:155-159
  zero_acc(acc);
  gemm_tile<TM, H, H>(s.buf0, LDA, W.w_kj, H, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = ty * 4 + p;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tx + 16 * q;
      float g = 0.f;
#pragma unroll
      for (int m = 0; m < BE; ++m) g = fmaf(__ldg(W.w_rbf2 + c * BE + m), s.r8[r * 8 + m], g);
      s.buf1[r * LDA + c] = __fmul_rn(swish(acc[p][q] + __ldg(W.b_kj + c)), g);
    }
  }
  __syncthreads();
  // x_kj = act(lin_down(x_kj))                                         spherenet.py:161
  float acc2[4][4];
  zero_acc(acc2);
  gemm_tile<TM, IE, H>(s.buf1, LDA, W.w_down, H, s.ws, acc2);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) s.buf0[(ty * 4 + p) * LDA + tx + 16 * q] = swish(acc2[p][q]);
  __syncthreads();
  tile_store<IE>(x_down + (size_t)e0 * IE, IE, s.buf0, LDA, rows);
}

// ------------------------------------------------------------------ update_e, part B
template <bool TORSION>
__global__ void __launch_bounds__(DT, 2)
sphere_update_e_b_kernel(const float* __restrict__ e1_in, const float* __restrict__ x_ji,
                         const float* __restrict__ x_down, const float* __restrict__ rbf0,
                         const float* __restrict__ sbf_p, const float* __restrict__ t_p, int ld_p,
                         const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                         const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ trip_ptr,
                         int n_edges, dig3d_update_e_weights W, float* __restrict__ e1_out,
                         float* __restrict__ v_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EdgeSmem& s = *reinterpret_cast<EdgeSmem*>(smem_raw);
  const int e0 = blockIdx.x * TM, rows = min(TM, n_edges - e0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = threadIdx.x; r < TM; r += DT) {
    s.src[r] = (r < rows) ? src[e0 + r] : -1;
    s.dst[r] = (r < rows) ? dst[e0 + r] : -1;
  }
  load_rbf_tile(s, rbf0, e0, rows);
  __syncthreads();
  // ---- triplet phase: m[e] = sum_{t in trip(e)} x_down[kj(t)] * lin_sbf2(sbf_p[t]) * lin_t2(t_p[t])
  //      spherenet.py:163-171.  Lane owns channels lane and lane+32; the second-stage basis weights
  //      of those two channels live in registers.
  {
    float ws2[2][BE], wt2[2][BE];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int m = 0; m < BE; ++m) {
        ws2[h][m] = __ldg(W.w_sbf2 + (lane + 32 * h) * BE + m);
        wt2[h][m] = TORSION ? __ldg(W.w_t2 + (lane + 32 * h) * BE + m) : 0.f;
      }
    for (int r = warp; r < TM; r += DT / 32) {
      float a0 = 0.f, a1 = 0.f;
      if (r < rows) {
        const int j = s.src[r], i = s.dst[r];
        const int base = row_ptr[j], d = row_ptr[j + 1] - base;
        int t = trip_ptr[e0 + r];
        for (int sl = 0; sl < d; ++sl) {
          const int kj = base + sl;
          if (src[kj] == i) continue;
          const float4* sp = reinterpret_cast<const float4*>(sbf_p + (size_t)t * ld_p);
          const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
          const float x0 = __ldg(x_down + (size_t)kj * IE + lane);
          const float x1 = __ldg(x_down + (size_t)kj * IE + lane + 32);
          const float sv[BE] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          float g0 = 0.f, g1 = 0.f;
#pragma unroll
          for (int m = 0; m < BE; ++m) { g0 = fmaf(ws2[0][m], sv[m], g0); g1 = fmaf(ws2[1][m], sv[m], g1); }
          float m0 = __fmul_rn(x0, g0), m1 = __fmul_rn(x1, g1);
          if (TORSION) {
            const float4* tp = reinterpret_cast<const float4*>(t_p + (size_t)t * ld_p);
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
            const float tv[BE] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            float h0 = 0.f, h1 = 0.f;
#pragma unroll
            for (int m = 0; m < BE; ++m) { h0 = fmaf(wt2[0][m], tv[m], h0); h1 = fmaf(wt2[1][m], tv[m], h1); }
            m0 = __fmul_rn(m0, h0); m1 = __fmul_rn(m1, h1);
          }
          a0 += m0; a1 += m1;
          ++t;
        }
      }
      s.buf1[r * LDA + lane] = a0;
      s.buf1[r * LDA + lane + 32] = a1;
    }
  }
  __syncthreads();
  float acc[4][8];
  // x_kj = act(lin_up(x_kj)); e1 = x_ji + x_kj                          spherenet.py:172-174
  zero_acc(acc);
  gemm_tile<TM, H, IE>(s.buf1, LDA, W.w_up, IE, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = ty * 4 + p;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tx + 16 * q;
      const float xj = (r < rows) ? __ldg(x_ji + (size_t)(e0 + r) * H + c) : 0.f;
      s.buf0[r * LDA + c] = xj + swish(acc[p][q]);
    }
  }
  __syncthreads();
  // residual layer: x + act(lin2(act(lin1(x))))                         spherenet.py:49-50
  auto residual = [&](float* x, float* tmp, const float* w1, const float* b1, const float* w2,
                      const float* b2) {
    zero_acc(acc);
    gemm_tile<TM, H, H>(x, LDA, w1, H, s.ws, acc);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = tx + 16 * q;
        tmp[(ty * 4 + p) * LDA + c] = swish(acc[p][q] + __ldg(b1 + c));
      }
    __syncthreads();
    zero_acc(acc);
    gemm_tile<TM, H, H>(tmp, LDA, w2, H, s.ws, acc);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = tx + 16 * q;
        x[(ty * 4 + p) * LDA + c] += swish(acc[p][q] + __ldg(b2 + c));
      }
    __syncthreads();
  };
  residual(s.buf0, s.buf1, W.w_res[0], W.b_res[0], W.w_res[1], W.b_res[1]);
  // e1 = act(lin(e1)) + x1                                              spherenet.py:177
  zero_acc(acc);
  gemm_tile<TM, H, H>(s.buf0, LDA, W.w_lin, H, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = ty * 4 + p;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tx + 16 * q;
      const float skip = (r < rows) ? __ldg(e1_in + (size_t)(e0 + r) * H + c) : 0.f;
      s.buf1[r * LDA + c] = swish(acc[p][q] + __ldg(W.b_lin + c)) + skip;
    }
  }
  __syncthreads();
  residual(s.buf1, s.buf0, W.w_res[2], W.b_res[2], W.w_res[3], W.b_res[3]);
  residual(s.buf1, s.buf0, W.w_res[4], W.b_res[4], W.w_res[5], W.b_res[5]);
  tile_store<H>(e1_out + (size_t)e0 * H, H, s.buf1, LDA, rows);
  // e2 = lin_rbf(rbf0) * e1, scattered to the target nodes              spherenet.py:180,211
  for (int id = threadIdx.x; id < TM * H; id += DT) {
    const int r = id / H, c = id % H;
    s.buf0[r * LDA + c] = __fmul_rn(rbf_dot(W.w_rbf, c, s.rbf + r * 8), s.buf1[r * LDA + c]);
  }
  __syncthreads();
  tile_segment_accumulate(s.buf0, LDA, s.dst, rows, v_in, H);
}

// ------------------------------------------------------------------ update_v (node MLP)
constexpr int VKC = 16;    // K chunk of the node MLP: 107 KB per CTA -> two CTAs per SM
struct NodeSmem {
  float buf0[TMV * LDV];
  float buf1[TMV * LDV];
  float ws[2 * OE * (VKC + 4)];
};

struct UpdateVBatch {
  dig3d_update_v_weights w[8];
};

__device__ __forceinline__ void sphere_update_v_body(const float* __restrict__ v_in, int n_nodes, int out_channels,
                                                     const dig3d_update_v_weights& W, float* __restrict__ v_out);

__global__ void __launch_bounds__(DT, 2)
sphere_update_v_kernel(const float* __restrict__ v_in, int n_nodes, int out_channels,
                       dig3d_update_v_weights W, float* __restrict__ v_out) {
  sphere_update_v_body(v_in, n_nodes, out_channels, W, v_out);
}

// All (num_layers + 1) node MLPs of a forward pass in ONE launch (grid.y = block): they only feed the
// readout, so they are deferred to the end where 5 x 72 CTAs fill the GPU instead of competing with
// the edge kernels of the next block.
__global__ void __launch_bounds__(DT, 2)
sphere_update_v_batched_kernel(const float* __restrict__ v_in_all, int n_nodes, int out_channels,
                               UpdateVBatch B, float* __restrict__ v_out_all) {
  const int b = blockIdx.y;
  sphere_update_v_body(v_in_all + (size_t)b * n_nodes * H, n_nodes, out_channels, B.w[b],
                       v_out_all + (size_t)b * n_nodes * out_channels);
}

__device__ __forceinline__ void sphere_update_v_body(const float* __restrict__ v_in, int n_nodes, int out_channels,
                                                     const dig3d_update_v_weights& W, float* __restrict__ v_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NodeSmem& s = *reinterpret_cast<NodeSmem*>(smem_raw);
  const int n0 = blockIdx.x * TMV, rows = min(TMV, n_nodes - n0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  tile_load<H>(s.buf0, LDV, v_in + (size_t)n0 * H, H, rows);
  for (int id = threadIdx.x; id < (TMV - rows) * H; id += DT)
    s.buf0[(rows + id / H) * LDV + id % H] = 0.f;
  __syncthreads();
  float acc[2][16];
  // v = lin_up(v)  (no activation)                                      spherenet.py:212
  zero_acc(acc);
  gemm_tile<TMV, OE, H, VKC>(s.buf0, LDV, W.w_up, H, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int c = tx + 16 * q;
      s.buf1[(ty * 2 + p) * LDV + c] = acc[p][q] + (W.b_up ? __ldg(W.b_up + c) : 0.f);
    }
  __syncthreads();
  float* cur = s.buf1;
  float* nxt = s.buf0;
  for (int l = 0; l < W.n_lins; ++l) {  // v = act(lin(v))                spherenet.py:213-214
    zero_acc(acc);
    gemm_tile<TMV, OE, OE, VKC>(cur, LDV, W.w_lins[l], OE, s.ws, acc);
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int c = tx + 16 * q;
        nxt[(ty * 2 + p) * LDV + c] = swish(acc[p][q] + __ldg(W.b_lins[l] + c));
      }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  // v = lin(v): [out_channels, O], no bias                              spherenet.py:215
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += DT / 32)
    for (int oc = 0; oc < out_channels; ++oc) {
      float part = 0.f;
      for (int c = lane; c < OE; c += 32) part = fmaf(cur[r * LDV + c], __ldg(W.w_out + (size_t)oc * OE + c), part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) v_out[(size_t)(n0 + r) * out_channels + oc] = part;
    }
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return DIG3D_ECUDA;
  }
  return DIG3D_OK;
}

}  // namespace dig3d

using namespace dig3d;

extern "C" {

int dig3d_sphere_init_e(const int64_t* z, const int32_t* src, const int32_t* dst, const float* rbf0,
                        int64_t n_edges, const dig3d_init_e_weights* w, float* e1, float* v_in,
                        void* stream) {
  DIG3D_REQUIRE(z && src && dst && rbf0 && w && e1 && v_in, "sphere_init_e: null pointer");
  DIG3D_REQUIRE(w->emb && w->w_rbf0 && w->b_rbf0 && w->w_lin && w->b_lin && w->w_rbf1,
                "sphere_init_e: null weight");
  if (n_edges == 0) return DIG3D_OK;
  int rc = set_smem(sphere_init_e_kernel, sizeof(EdgeSmem));
  if (rc) return rc;
  sphere_init_e_kernel<<<ceil_div(n_edges, TM), DT, sizeof(EdgeSmem), (cudaStream_t)stream>>>(
      z, src, dst, rbf0, (int)n_edges, *w, e1, v_in);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_e_a(const float* e1, const float* rbf0, int64_t n_edges,
                            const dig3d_update_e_weights* w, float* x_ji, float* x_down, void* stream) {
  DIG3D_REQUIRE(e1 && rbf0 && w && x_ji && x_down, "sphere_update_e_a: null pointer");
  if (n_edges == 0) return DIG3D_OK;
  int rc = set_smem(sphere_update_e_a_kernel, sizeof(EdgeSmem));
  if (rc) return rc;
  sphere_update_e_a_kernel<<<ceil_div(n_edges, TM), DT, sizeof(EdgeSmem), (cudaStream_t)stream>>>(
      e1, rbf0, (int)n_edges, *w, x_ji, x_down);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_e_b(const float* e1_in, const float* x_ji, const float* x_down, const float* rbf0,
                            const float* sbf_p, const float* t_p, int32_t ld_p, const int32_t* src,
                            const int32_t* dst, const int32_t* row_ptr, const int32_t* trip_ptr,
                            int64_t n_edges, const dig3d_update_e_weights* w, float* e1_out, float* v_in,
                            void* stream) {
  DIG3D_REQUIRE(e1_in && x_ji && x_down && rbf0 && sbf_p && src && dst && row_ptr && trip_ptr && w && e1_out &&
                    v_in, "sphere_update_e_b: null pointer");
  DIG3D_REQUIRE((t_p != nullptr) == (w->w_t2 != nullptr), "sphere_update_e_b: t_p and w_t2 must agree");
  DIG3D_REQUIRE(ld_p % 4 == 0, "sphere_update_e_b: ld_p must be a multiple of 4");
  if (n_edges == 0) return DIG3D_OK;
  const int grid = ceil_div(n_edges, TM);
  if (t_p) {
    int rc = set_smem(sphere_update_e_b_kernel<true>, sizeof(EdgeSmem));
    if (rc) return rc;
    sphere_update_e_b_kernel<true><<<grid, DT, sizeof(EdgeSmem), (cudaStream_t)stream>>>(
        e1_in, x_ji, x_down, rbf0, sbf_p, t_p, ld_p, src, dst, row_ptr, trip_ptr, (int)n_edges, *w, e1_out,
        v_in);
  } else {
    int rc = set_smem(sphere_update_e_b_kernel<false>, sizeof(EdgeSmem));
    if (rc) return rc;
    sphere_update_e_b_kernel<false><<<grid, DT, sizeof(EdgeSmem), (cudaStream_t)stream>>>(
        e1_in, x_ji, x_down, rbf0, sbf_p, t_p, ld_p, src, dst, row_ptr, trip_ptr, (int)n_edges, *w, e1_out,
        v_in);
  }
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_v_batched(const float* v_in_all, int64_t n_nodes, int32_t n_blocks, int32_t out_channels,
                                  const dig3d_update_v_weights* w, float* v_out_all, void* stream) {
  DIG3D_REQUIRE(v_in_all && w && v_out_all && out_channels > 0, "sphere_update_v_batched: bad arguments");
  DIG3D_REQUIRE(n_blocks >= 1 && n_blocks <= 8, "sphere_update_v_batched: n_blocks=%d outside [1,8]", n_blocks);
  if (n_nodes == 0) return DIG3D_OK;
  UpdateVBatch B;
  for (int b = 0; b < n_blocks; ++b) {
    DIG3D_REQUIRE(w[b].n_lins >= 0 && w[b].n_lins <= 8, "sphere_update_v_batched: n_lins outside [0,8]");
    B.w[b] = w[b];
  }
  int rc = set_smem(sphere_update_v_batched_kernel, sizeof(NodeSmem));
  if (rc) return rc;
  dim3 grid(ceil_div(n_nodes, TMV), n_blocks);
  sphere_update_v_batched_kernel<<<grid, DT, sizeof(NodeSmem), (cudaStream_t)stream>>>(
      v_in_all, (int)n_nodes, out_channels, B, v_out_all);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_v(const float* v_in, int64_t n_nodes, int32_t out_channels,
                          const dig3d_update_v_weights* w, float* v_out, void* stream) {
  DIG3D_REQUIRE(v_in && w && v_out && out_channels > 0, "sphere_update_v: bad arguments");
  DIG3D_REQUIRE(w->n_lins >= 0 && w->n_lins <= 8, "sphere_update_v: n_lins=%d outside [0,8]", w->n_lins);
  if (n_nodes == 0) return DIG3D_OK;
  int rc = set_smem(sphere_update_v_kernel, sizeof(NodeSmem));
  if (rc) return rc;
  sphere_update_v_kernel<<<ceil_div(n_nodes, TMV), DT, sizeof(NodeSmem), (cudaStream_t)stream>>>(
      v_in, (int)n_nodes, out_channels, *w, v_out);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

}  // extern "C"
