// SchNet interaction (continuous-filter convolution) on sm_100a, fp32.
//
//   emb (Gaussian smearing)   schnet.py:85-94
//   update_e.forward          schnet.py:29-35    W = mlp(dist_emb) * C(d);  e = lin(v)[j] * W
//   update_v.forward          schnet.py:53-59    v + lin2(ssp(lin1(scatter(e, i))))
//   update_u.forward          schnet.py:77-82
//
// Fusion: one kernel per interaction evaluates, for a tile of 64 target-sorted edges, the Gaussian
// expansion, the two-layer filter MLP, the cosine cutoff, the gather of lin(v)[j], the Hadamard
// product and the segmented sum over the target node -- the [E, G] expansion, the filter [E, F] and
// the messages [E, F] never touch HBM.
#include "dense.cuh"

namespace dig3d {

constexpr int SG = 64;  // num_gaussians padded to a multiple of the K chunk (weights are zero padded)

__device__ __forceinline__ float ssp(float x) {
  // F.softplus(x) - log(2)  (beta = 1, threshold = 20)            schnet.py:97-103
  const float sp = (x > 20.0f) ? x : log1pf(expf(x));
  return __fsub_rn(sp, 0.693147182464599609375f);
}

template <int F>
struct SchnetEdgeSmem {
  static constexpr int LDG_ = SG + 4, LDF = F + 4;
  float g[64 * LDG_];       // Gaussian expansion, later reused for the second activation
  float h[64 * LDF];
  float ws[2 * F * LDW];
  float dist[64];
  int src[64];
  int dst[64];
};

// out[N, F] += segment_sum_i( lin(v)[j] * mlp(gauss(d)) * C(d) )
template <int F>
__global__ void __launch_bounds__(DT, 2)
schnet_cfconv_kernel(const float* __restrict__ dist, const int32_t* __restrict__ src,
                     const int32_t* __restrict__ dst, int n_edges, const float* __restrict__ vlin,
                     const float* __restrict__ offset, int n_gauss, float coeff, float pi_over_cutoff_a,
                     float inv_cutoff, const float* __restrict__ w0 /*[F, SG] zero padded*/,
                     const float* __restrict__ b0, const float* __restrict__ w2, const float* __restrict__ b2,
                     float* __restrict__ out) {
  using S = SchnetEdgeSmem<F>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& s = *reinterpret_cast<S*>(smem_raw);
  constexpr int NQ = F / 16;
  const int e0 = blockIdx.x * 64, rows = min(64, n_edges - e0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (int r = threadIdx.x; r < 64; r += DT) {
    s.src[r] = (r < rows) ? src[e0 + r] : -1;
    s.dst[r] = (r < rows) ? dst[e0 + r] : -1;
    s.dist[r] = (r < rows) ? dist[e0 + r] : 0.f;
  }
  __syncthreads();
  // exp(coeff * (d - mu_g)^2)                                        schnet.py:92-94
  for (int id = threadIdx.x; id < 64 * SG; id += DT) {
    const int r = id / SG, gg = id % SG;
    float v = 0.f;
    if (gg < n_gauss) {
      const float t = __fsub_rn(s.dist[r], __ldg(offset + gg));
      v = expf(__fmul_rn(coeff, __fmul_rn(t, t)));
    }
    s.g[r * S::LDG_ + gg] = v;
  }
  __syncthreads();
  float acc[4][NQ];
  zero_acc(acc);
  gemm_tile<64, F, SG>(s.g, S::LDG_, w0, SG, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = tx + 16 * q;
      s.h[(ty * 4 + p) * S::LDF + c] = ssp(acc[p][q] + __ldg(b0 + c));
    }
  __syncthreads();
  zero_acc(acc);
  gemm_tile<64, F, F>(s.h, S::LDF, w2, F, s.ws, acc);
  // W = mlp(.) * C,  e = lin(v)[j] * W                                 schnet.py:31-34
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = ty * 4 + p;
    // C = 0.5 * (cos(d * pi / cutoff) + 1): (d*pi) then *(1/cutoff) as ATen-CUDA evaluates it
    const float cc = __fmul_rn(0.5f, __fadd_rn(cosf(__fmul_rn(__fmul_rn(s.dist[r], pi_over_cutoff_a), inv_cutoff)), 1.0f));
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = tx + 16 * q;
      const float wv = __fmul_rn(acc[p][q] + __ldg(b2 + c), cc);
      const float vj = (r < rows) ? __ldg(vlin + (size_t)s.src[r] * F + c) : 0.f;
      s.h[r * S::LDF + c] = __fmul_rn(vj, wv);    // s.h (GEMM 2's input) is free again
    }
  }
  __syncthreads();
  tile_segment_accumulate(s.h, S::LDF, s.dst, rows, out, F);
}

// Node-level linears: y = [x +] act?(x W^T + b) chains used by SchNet (tile of 64 nodes).
//   mode 0: y = x W^T                     (update_e.lin, no bias)       schnet.py:33
//   mode 1: y = v + lin2(ssp(lin1(a)))    (update_v)                    schnet.py:56-59
//   mode 2: y = lin2(ssp(lin1(v)))        (update_u before the scatter) schnet.py:78-80
template <int KI, int KM>
struct SchnetNodeSmem {
  float a[64 * (KI + 4)];
  float m[64 * (KM + 4)];
  float ws[2 * ((KI > KM ? KI : KM)) * LDW];
};

template <int KI, int NO>
__global__ void __launch_bounds__(DT, 2)
schnet_linear_kernel(const float* __restrict__ x, int n_rows, const float* __restrict__ w, float* __restrict__ y) {
  using S = SchnetNodeSmem<KI, NO>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& s = *reinterpret_cast<S*>(smem_raw);
  const int r0 = blockIdx.x * 64, rows = min(64, n_rows - r0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  tile_load<KI>(s.a, KI + 4, x + (size_t)r0 * KI, KI, rows);
  for (int id = threadIdx.x; id < (64 - rows) * KI; id += DT) s.a[(rows + id / KI) * (KI + 4) + id % KI] = 0.f;
  __syncthreads();
  float acc[4][NO / 16];
  zero_acc(acc);
  gemm_tile<64, NO, KI>(s.a, KI + 4, w, KI, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < NO / 16; ++q) s.m[(ty * 4 + p) * (NO + 4) + tx + 16 * q] = acc[p][q];
  __syncthreads();
  tile_store<NO>(y + (size_t)r0 * NO, NO, s.m, NO + 4, rows);
}

// y = [res +] lin2(ssp(lin1(a)))   lin1: [KM, KI] + b1, lin2: [NO, KM] + b2
template <int KI, int KM, int NO>
__global__ void __launch_bounds__(DT, 2)
schnet_mlp2_kernel(const float* __restrict__ a, int n_rows, const float* __restrict__ w1,
                   const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                   const float* __restrict__ res, float* __restrict__ y) {
  using S = SchnetNodeSmem<KI, (KM > NO ? KM : NO)>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& s = *reinterpret_cast<S*>(smem_raw);
  constexpr int LDM = (KM > NO ? KM : NO) + 4;
  const int r0 = blockIdx.x * 64, rows = min(64, n_rows - r0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  tile_load<KI>(s.a, KI + 4, a + (size_t)r0 * KI, KI, rows);
  for (int id = threadIdx.x; id < (64 - rows) * KI; id += DT) s.a[(rows + id / KI) * (KI + 4) + id % KI] = 0.f;
  __syncthreads();
  float acc[4][KM / 16];
  zero_acc(acc);
  gemm_tile<64, KM, KI>(s.a, KI + 4, w1, KI, s.ws, acc);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < KM / 16; ++q) {
      const int c = tx + 16 * q;
      s.m[(ty * 4 + p) * LDM + c] = ssp(acc[p][q] + __ldg(b1 + c));
    }
  __syncthreads();
  float acc2[4][NO / 16];
  zero_acc(acc2);
  gemm_tile<64, NO, KM>(s.m, LDM, w2, KM, s.ws, acc2);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = ty * 4 + p;
#pragma unroll
    for (int q = 0; q < NO / 16; ++q) {
      const int c = tx + 16 * q;
      float v = acc2[p][q] + __ldg(b2 + c);
      if (res && r < rows) v = __ldg(res + (size_t)(r0 + r) * NO + c) + v;
      s.a[r * (KI + 4) + c] = v;   // NO <= KI in every instantiation
    }
  }
  __syncthreads();
  tile_store<NO>(y + (size_t)r0 * NO, NO, s.a, KI + 4, rows);
}

// update_u before the scatter: node_out[n] = lin2(ssp(lin1(v[n])))  (tiny: one warp per node)  schnet.py:78-80
constexpr int RO_MAXH2 = 128;
__global__ void __launch_bounds__(256)
schnet_readout_kernel(const float* __restrict__ v, int n_nodes, int hidden, const float* __restrict__ w1,
                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                      int out_channels, float* __restrict__ node_out) {
  __shared__ float hbuf[8][RO_MAXH2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + w;
  if (n >= n_nodes) return;
  const int h2 = hidden / 2;
  const float* vr = v + (size_t)n * hidden;
  for (int u = lane; u < h2; u += 32) {
    float acc = 0.f;
    for (int k = 0; k < hidden; ++k) acc = fmaf(__ldg(vr + k), __ldg(w1 + (size_t)u * hidden + k), acc);
    hbuf[w][u] = ssp(acc + __ldg(b1 + u));
  }
  __syncwarp();
  for (int oc = 0; oc < out_channels; ++oc) {
    float part = 0.f;
    for (int u = lane; u < h2; u += 32) part = fmaf(hbuf[w][u], __ldg(w2 + (size_t)oc * h2 + u), part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) node_out[(size_t)n * out_channels + oc] = part + __ldg(b2 + oc);
  }
}

template <class K>
static int set_smem_attr(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return DIG3D_ECUDA;
  }
  return DIG3D_OK;
}

template <int HF>
static int schnet_block(const float* v, int64_t n_nodes, const float* dist, const int32_t* src, const int32_t* dst,
                        int64_t n_edges, const float* offset, int n_gauss, double coeff, double cutoff,
                        const dig3d_schnet_block_weights* w, float* vlin, float* agg, float* v_out,
                        cudaStream_t st) {
  // H == F == HF
  int rc;
  {
    auto k = schnet_linear_kernel<HF, HF>;
    const size_t sm = sizeof(SchnetNodeSmem<HF, HF>);
    if ((rc = set_smem_attr(k, sm))) return rc;
    k<<<ceil_div(n_nodes, 64), DT, sm, st>>>(v, (int)n_nodes, w->w_lin, vlin);
  }
  if (n_edges) {
    auto k = schnet_cfconv_kernel<HF>;
    const size_t sm = sizeof(SchnetEdgeSmem<HF>);
    if ((rc = set_smem_attr(k, sm))) return rc;
    const float pi_f = 3.14159274101257324f;
    k<<<ceil_div(n_edges, 64), DT, sm, st>>>(dist, src, dst, (int)n_edges, vlin, offset, n_gauss, (float)coeff,
                                             pi_f, 1.0f / (float)cutoff, w->w_mlp0, w->b_mlp0, w->w_mlp2,
                                             w->b_mlp2, agg);
  }
  {
    auto k = schnet_mlp2_kernel<HF, HF, HF>;
    const size_t sm = sizeof(SchnetNodeSmem<HF, HF>);
    if ((rc = set_smem_attr(k, sm))) return rc;
    k<<<ceil_div(n_nodes, 64), DT, sm, st>>>(agg, (int)n_nodes, w->w_v1, w->b_v1, w->w_v2, w->b_v2, v, v_out);
  }
  return DIG3D_OK;
}

}  // namespace dig3d

using namespace dig3d;

extern "C" {

int dig3d_schnet_block(const float* v, int64_t n_nodes, const float* dist, const int32_t* src,
                       const int32_t* dst, int64_t n_edges, const float* offset, int32_t n_gauss, double coeff,
                       double cutoff, int32_t hidden, int32_t filters, const dig3d_schnet_block_weights* w,
                       float* vlin, float* agg, float* v_out, void* stream) {
  DIG3D_REQUIRE(v && dist && src && dst && offset && w && vlin && agg && v_out, "schnet_block: null pointer");
  DIG3D_REQUIRE(n_gauss >= 1 && n_gauss <= SG, "schnet_block: num_gaussians=%d outside [1,%d]", n_gauss, SG);
  DIG3D_REQUIRE(hidden == filters && (hidden == 32 || hidden == 64 || hidden == 128),
                "schnet_block: compiled for hidden == num_filters in {32, 64, 128}, got %d/%d", hidden, filters);
  if (n_nodes == 0) return DIG3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  switch (hidden) {
    case 32: rc = schnet_block<32>(v, n_nodes, dist, src, dst, n_edges, offset, n_gauss, coeff, cutoff, w, vlin, agg, v_out, st); break;
    case 64: rc = schnet_block<64>(v, n_nodes, dist, src, dst, n_edges, offset, n_gauss, coeff, cutoff, w, vlin, agg, v_out, st); break;
    default: rc = schnet_block<128>(v, n_nodes, dist, src, dst, n_edges, offset, n_gauss, coeff, cutoff, w, vlin, agg, v_out, st); break;
  }
  if (rc) return rc;
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_schnet_readout(const float* v, int64_t n_nodes, int32_t hidden, const float* w1, const float* b1,
                         const float* w2, const float* b2, int32_t out_channels, float* node_out, void* stream) {
  DIG3D_REQUIRE(v && w1 && b1 && w2 && b2 && node_out, "schnet_readout: null pointer");
  DIG3D_REQUIRE(hidden >= 2 && hidden / 2 <= RO_MAXH2, "schnet_readout: hidden=%d unsupported", hidden);
  if (n_nodes == 0) return DIG3D_OK;
  schnet_readout_kernel<<<ceil_div(n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
      v, (int)n_nodes, hidden, w1, b1, w2, b2, out_channels, node_out);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

}  // extern "C"
