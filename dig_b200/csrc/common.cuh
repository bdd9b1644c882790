// Shared helpers for libdig3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dig3d.h"

namespace dig3d {

void set_error(const char* fmt, ...);

#define DIG3D_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      dig3d::set_error(__VA_ARGS__);             \
      return DIG3D_EINVAL;                       \
    }                                            \
  } while (0)

#define DIG3D_LAUNCH_CHECK()                                                    \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      dig3d::set_error("%s:%d CUDA launch failed: %s", __FILE__, __LINE__,      \
                       cudaGetErrorString(e__));                                \
      return DIG3D_ECUDA;                                                       \
    }                                                                           \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- ATen-CUDA rounding forms (verified against the sm_100 SASS of
//      at::native::cross_kernel<float> and against ATen/native/cuda/Reduce.cuh) -------------
struct f3 {
  float x, y, z;
};

__device__ __forceinline__ f3 sub3(const f3 a, const f3 b) {
  return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)};
}
__device__ __forceinline__ f3 mul3(const f3 a, const f3 b) {
  return {__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)};
}
// at::native::cross_kernel: the positive product is fused, the subtracted one rounded first.
__device__ __forceinline__ f3 cross_aten(const f3 a, const f3 b) {
  return {__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)),
          __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
          __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x))};
}
// sum over a contiguous innermost dim of size 3: the reduce kernel maps it on 2 lanes,
// lane 0 adds elements 0 and 2, then the shuffle adds lane 1's element 1.
__device__ __forceinline__ float sum3_aten(const f3 v) {
  return __fadd_rn(__fadd_rn(v.x, v.z), v.y);
}
// (v*v).sum().sqrt()  ==  v.norm()  (NormTwoOps: fma(v,v,0) == rn(v*v), same tree)
__device__ __forceinline__ float norm3_aten(const f3 v) {
  return __fsqrt_rn(sum3_aten(mul3(v, v)));
}
__device__ __forceinline__ f3 load3(const float* __restrict__ p, int n) {
  return {__ldg(p + 3 * n), __ldg(p + 3 * n + 1), __ldg(p + 3 * n + 2)};
}

// ---- packed fp32 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 on 64-bit register pairs) ----------------------------
// Each half is an ordinary round-to-nearest fp32 operation, so a packed chain is bit-identical to the two scalar
// chains it replaces; the FP32 pipe issues one packed instruction in the slot of one scalar FFMA
// (tools/ffma_rate.cu, profiles/r02_ffma2_rate.txt).
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<const unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(const float2 a, const float2 b) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float swish(float x) {
  // x * sigmoid(x); ATen sigmoid = 1 / (1 + exp(-x)) in fp32
  return __fmul_rn(x, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))));
}

}  // namespace dig3d
