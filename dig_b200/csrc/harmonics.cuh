// Real spherical harmonics by recurrence (sm_100a), for the FUSED projection kernel only.
//
// The reference builds its harmonics symbolically from exactly these recurrences (spherenet/features.py:74-96
// associated Legendre P_l^m(z) without the (1 - z^2)^(m/2) factor, :113-126 the polynomials S_m, C_m in
// x = sin(theta) cos(phi), y = sin(theta) sin(phi), :99-148 the assembly) and then lets sympy.simplify rewrite every
// entry into a closed form, which it lambdifies.  The bit-exact basis kernels (dig3d_triplet_basis, the generated
// headers) evaluate those closed forms node by node with the reference's rounding -- ~1200 instructions per triplet,
// 22 sinf / cosf and 8 powf calls.  The fused projection never exposes the basis itself (only its contraction with
// lin_sbf1 / lin_t1, compared with the oracle at 2e-6), so it evaluates the SAME functions from the recurrences:
// two sincosf and ~250 multiply-adds.  tests/test_basis.py pins the numpy twin of this file
// (dig_b200/basis.py harmonics_recurrence) against the reference's closed forms.
#pragma once
#include "common.cuh"

namespace dig3d {
namespace hrec {

constexpr double csqrt(double x) {
  double r = x > 1.0 ? x : 1.0;
  for (int i = 0; i < 80; ++i) r = 0.5 * (r + x / r);
  return r;
}
constexpr double cfact(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}
constexpr double kPi = 3.14159265358979323846;
// features.py:69-71 sph_harm_prefactor(l, m); the m != 0 entries carry an extra sqrt(2) (features.py:139-146)
constexpr double norm(int l, int m) {
  return csqrt((2 * l + 1) * cfact(l - m) / (4.0 * kPi * cfact(l + m))) * (m == 0 ? 1.0 : csqrt(2.0));
}
// every coefficient of the recurrences for l, m < L, evaluated by the compiler (the loops below are fully unrolled, so
// each use is an immediate operand)
template <int L>
struct Table {
  float n[L][L], a[L][L], b[L][L];      // norm; P_l^m = a z P_(l-1)^m - b P_(l-2)^m
  constexpr Table() : n{}, a{}, b{} {
    for (int l = 0; l < L; ++l)
      for (int m = 0; m <= l; ++m) {
        n[l][m] = (float)norm(l, m);
        a[l][m] = l > m ? (float)((2.0 * l - 1.0) / (l - m)) : 0.f;
        b[l][m] = l > m ? (float)((double)(l + m - 1) / (l - m)) : 0.f;
      }
  }
};

}  // namespace hrec

// Zonal harmonics Y_l^0(theta), l < L.
template <int L>
__device__ __forceinline__ void yl0_recurrence(const float theta, float (&y0)[L]) {
  constexpr hrec::Table<L> T{};
  const float z = cosf(theta);
  float p2 = 1.f, p1 = z;
  y0[0] = T.n[0][0];
  if (L > 1) y0[1] = T.n[1][0] * z;
#pragma unroll
  for (int l = 2; l < L; ++l) {
    const float p = (T.a[l][0] * z) * p1 - T.b[l][0] * p2;
    y0[l] = T.n[l][0] * p;
    p2 = p1; p1 = p;
  }
}

// All L*L harmonics in the reference's flat order (per l: m = 0, +1 .. +l, -l .. -1; entry l*l + k) plus the zonal
// ones again in y0 (the angle basis uses them with its own weights).
template <int L>
__device__ __forceinline__ void ylm_recurrence(const float theta, const float phi, float (&y)[L * L], float (&y0)[L]) {
  constexpr hrec::Table<L> T{};
  float st, ct, sp, cp;
  sincosf(theta, &st, &ct);
  sincosf(phi, &sp, &cp);
  const float x = st * cp, yy = st * sp, z = ct;
  // m = 0 column
  {
    float p2 = 1.f, p1 = z;
    y0[0] = y[0] = T.n[0][0];
    if (L > 1) y0[1] = y[1] = T.n[1][0] * z;
#pragma unroll
    for (int l = 2; l < L; ++l) {
      const float p = (T.a[l][0] * z) * p1 - T.b[l][0] * p2;
      y0[l] = y[l * l] = T.n[l][0] * p;
      p2 = p1; p1 = p;
    }
  }
  // m >= 1: S_m, C_m by the angle-addition recurrence, P_l^m upwards in l from P_m^m = (1 - 2m) P_(m-1)^(m-1)
  float s = 0.f, c = 1.f, pmm = 1.f;
#pragma unroll
  for (int m = 1; m < L; ++m) {
    const float s_new = x * s + yy * c, c_new = x * c - yy * s;
    s = s_new; c = c_new;
    pmm = (float)(1 - 2 * m) * pmm;
    float p2 = 0.f, p1 = pmm;
#pragma unroll
    for (int l = m; l < L; ++l) {
      float p;
      if (l == m) p = pmm;
      else if (l == m + 1) p = ((float)(2 * m + 1) * z) * pmm;
      else p = (T.a[l][m] * z) * p1 - T.b[l][m] * p2;
      const float np = T.n[l][m] * p;
      y[l * l + m] = np * c;                    // +m
      y[l * l + 2 * l + 1 - m] = np * s;        // -m (python index -m of a 2l+1 list)
      p2 = p1; p1 = p;
    }
  }
}

}  // namespace dig3d
