// ProNet edge geometry + basis features (sm_100a).      reference dig/threedgraph/method/pronet/pronet.py:352-449,
//                                                       pronet/features.py:253-344 (= the ComENet closed forms, nr = 6)
//
// One thread per edge (j -> i) of the C-alpha radius graph:
//   dist, theta, phi from the SEQUENCE neighbours (i-1) mod N, (i+1) mod N of the target residue       pronet.py:383-402
//   level 0 (aminoacid): tau from the sequence neighbours of i and j                                    pronet.py:430-449
//   level 1 (backbone / allatom): the three Euler angles between the local frames built from N, CA, C  pronet.py:404-426
//   feature0[E, 24] = d_theta_phi_emb(dist, theta, phi);  feature1[E, 12 | 36] = d_angle_emb(dist, tau | angle1..3)
//   pos_emb[E, 16] = [cos((j - i) f_k), sin((j - i) f_k)],  f_k = exp(-2k ln(1e4) / 16)                  pronet.py:352-362
// Same ATen-CUDA rounding helpers as the other geometry kernels (cross / 3-sum / norm), so the angles match the
// reference op for op.
#include "common.cuh"
#include "generated/basis_gemnet_2_6.cuh"

namespace dig3d {

constexpr int PN_NR = 6, PN_NS = 2;
constexpr int PN_F0 = PN_NR * PN_NS * PN_NS;   // 24
constexpr int PN_F1 = PN_NR * PN_NS;           // 12 per angle

__device__ __forceinline__ float dot3_aten(const f3 a, const f3 b) { return sum3_aten(mul3(a, b)); }

__device__ __forceinline__ void pn_angle_features(const float (&rb)[PN_NR * PN_NS], float ang, float* __restrict__ out) {
  float y0[PN_NS];
  basis_gemnet_2_6::yl0(ang, y0);
#pragma unroll
  for (int l = 0; l < PN_NS; ++l)
#pragma unroll
    for (int r = 0; r < PN_NR; ++r) out[l * PN_NR + r] = __fmul_rn(rb[l * PN_NR + r], y0[l]);
}

__global__ void pronet_edge_features_kernel(const float* __restrict__ pos, const float* __restrict__ pos_n,
                                            const float* __restrict__ pos_c, const int32_t* __restrict__ src,
                                            const int32_t* __restrict__ dst, int n_edges, int n_nodes, int level,
                                            float inv_cutoff, int num_pos_emb, float* __restrict__ dist_out,
                                            float* __restrict__ f0, float* __restrict__ f1, float* __restrict__ pe,
                                            float* __restrict__ angles) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int j = src[e], i = dst[e];
  const f3 pi_ = load3(pos, i), pj = load3(pos, j);
  const float dist = norm3_aten(sub3(pi_, pj));                            // (pos[i] - pos[j]).norm()
  const int refi0 = (i - 1 + n_nodes) % n_nodes, refi1 = (i + 1) % n_nodes;
  const f3 vji = sub3(pj, pi_);                                            // pos[j] - pos[i]
  const f3 v0 = sub3(load3(pos, refi0), pi_), v1 = sub3(load3(pos, refi1), pi_);
  const float theta = atan2f(norm3_aten(cross_aten(vji, v0)), dot3_aten(vji, v0));
  const f3 plane1 = cross_aten(v0, v1), plane2 = cross_aten(v0, vji);
  const float phi = atan2f(__fdiv_rn(dot3_aten(cross_aten(plane1, plane2), v0), norm3_aten(v0)), dot3_aten(plane1, plane2));
  float ang[3];
  int n_ang;
  if (level == 0) {
    int refi = refi0, refj = (j - 1 + n_nodes) % n_nodes;
    const int refj1 = (j + 1) % n_nodes;
    if (refi0 == j) refi = refi1;
    if (refj == i) refj = refj1;
    const f3 q1 = cross_aten(vji, sub3(load3(pos, refi), pi_));
    const f3 q2 = cross_aten(vji, sub3(load3(pos, refj), pj));
    ang[0] = atan2f(__fdiv_rn(dot3_aten(cross_aten(q1, q2), vji), dist), dot3_aten(q1, q2));
    n_ang = 1;
  } else {
    const f3 o1x = sub3(load3(pos_n, i), pi_);
    const f3 o1z = cross_aten(o1x, cross_aten(o1x, sub3(load3(pos_c, i), pi_)));
    const float o1l = __fadd_rn(norm3_aten(o1z), 1e-7f);
    const f3 o2x = sub3(load3(pos_n, j), pj);
    const f3 o2z = cross_aten(o2x, cross_aten(o2x, sub3(load3(pos_c, j), pj)));
    const float o2l = __fadd_rn(norm3_aten(o2z), 1e-7f);
    const f3 nn = cross_aten(o1z, o2z);
    ang[0] = atan2f(__fdiv_rn(dot3_aten(cross_aten(o1x, nn), o1z), o1l), dot3_aten(o1x, nn));
    ang[1] = atan2f(norm3_aten(nn), dot3_aten(o1z, o2z));
    ang[2] = atan2f(__fdiv_rn(dot3_aten(cross_aten(nn, o2x), o2z), o2l), dot3_aten(nn, o2x));
    n_ang = 3;
  }
  if (dist_out) dist_out[e] = dist;
  if (angles) {
    angles[(size_t)e * 5] = theta;
    angles[(size_t)e * 5 + 1] = phi;
    for (int a = 0; a < 3; ++a) angles[(size_t)e * 5 + 2 + a] = a < n_ang ? ang[a] : 0.f;
  }
  // basis features
  const float x = __fmul_rn(dist, inv_cutoff);
  float rb[PN_NR * PN_NS], ylm[PN_NS * PN_NS];
  basis_gemnet_2_6::bessel(x, rb);
  basis_gemnet_2_6::ylm(theta, phi, ylm);
#pragma unroll
  for (int h = 0; h < PN_NS * PN_NS; ++h)
#pragma unroll
    for (int r = 0; r < PN_NR; ++r) f0[(size_t)e * PN_F0 + h * PN_NR + r] = __fmul_rn(rb[(h == 0 ? 0 : 1) * PN_NR + r], ylm[h]);
  for (int a = 0; a < n_ang; ++a) pn_angle_features(rb, ang[a], f1 + (size_t)e * (PN_F1 * n_ang) + a * PN_F1);
  // positional embedding: d = j - i; angles = d * exp(arange(0, P, 2) * -(ln 1e4 / P)) evaluated like torch does
  const float d = (float)(j - i);
  const int half = num_pos_emb / 2;
  const float scale = (float)(-(9.210340371976184 / (double)num_pos_emb));   // -(np.log(10000.0) / P) as fp32 scalar
  for (int k = 0; k < half; ++k) {
    const float fr = expf(__fmul_rn((float)(2 * k), scale));
    const float a = __fmul_rn(d, fr);
    pe[(size_t)e * num_pos_emb + k] = cosf(a);
    pe[(size_t)e * num_pos_emb + half + k] = sinf(a);
  }
}

}  // namespace dig3d

using namespace dig3d;

extern "C" int dig3d_pronet_edge_features(const float* pos_ca, const float* pos_n, const float* pos_c, const int32_t* src,
                                          const int32_t* dst, int64_t n_edges, int64_t n_nodes, int32_t level,
                                          double cutoff, int32_t num_pos_emb, float* dist, float* feature0,
                                          float* feature1, float* pos_emb, float* angles, void* stream) {
  DIG3D_REQUIRE(pos_ca && src && dst && feature0 && feature1 && pos_emb, "pronet_edge_features: null pointer");
  DIG3D_REQUIRE(level == 0 || (pos_n && pos_c), "pronet_edge_features: backbone / allatom levels need coords_n / coords_c");
  DIG3D_REQUIRE(level == 0 || level == 1, "pronet_edge_features: level must be 0 (aminoacid) or 1 (backbone / allatom)");
  DIG3D_REQUIRE(num_pos_emb > 0 && num_pos_emb % 2 == 0, "pronet_edge_features: num_pos_emb must be even");
  if (n_edges == 0) return DIG3D_OK;
  pronet_edge_features_kernel<<<ceil_div(n_edges, 128), 128, 0, (cudaStream_t)stream>>>(
      pos_ca, pos_n, pos_c, src, dst, (int)n_edges, (int)n_nodes, level, 1.0f / (float)cutoff, num_pos_emb, dist,
      feature0, feature1, pos_emb, angles);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}
