// CTA-level fp32 dense tile engine used by the fused interaction-block kernels.
//
//   acc[TM x NOUT] += A[TM x K] (shared memory activations) * W[NOUT x K]^T (nn.Linear layout,
//   streamed from global/L2 in K-chunks through a cp.async double buffer).
//
// Exact fp32 FFMA accumulation (k ascending) -- this is the parity-first implementation of the
// edge-MLP contraction; the reference runs these as cuBLAS SGEMM with TF32 disabled.
//
// 256 threads.  Thread (ty = tid/16, tx = tid%16) owns rows ty*RP + p (RP = TM/16) and the strided
// columns tx + 16*q (q < NOUT/16): with a KC+4 row stride for the staged weights the eight 16-byte
// weight reads of a quarter-warp fall in eight different bank groups, and activation reads are
// warp broadcasts.
#pragma once
#include "common.cuh"

namespace dig3d {

constexpr int DT = 256;   // threads per CTA
constexpr int KC = 32;    // K chunk
constexpr int LDW = KC + 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NOUT, int KCH>
__device__ __forceinline__ void stage_weights(const float* __restrict__ Wg, int ldwg, int k0, float* Ws) {
  // NOUT rows x KCH floats; KCH/4 16-byte segments per row
  constexpr int SEGS = NOUT * (KCH / 4), PER = KCH / 4, LD = KCH + 4;
#pragma unroll
  for (int it = 0; it < SEGS / DT; ++it) {
    const int id = threadIdx.x + it * DT;
    const int o = id / PER, part = id % PER;
    cp_async16(Ws + o * LD + part * 4, Wg + (size_t)o * ldwg + k0 + part * 4);
  }
}

// Ws must hold 2 * NOUT * (KCH + 4) floats.  All threads must call; ends with __syncthreads().
// KCH = K chunk staged per step (32 by default; 16 halves the staging footprint, used where it buys a
// second resident CTA per SM).
template <int TM, int NOUT, int K, int KCH = KC>
__device__ __forceinline__ void gemm_tile(const float* As, int lda, const float* __restrict__ Wg, int ldwg,
                                          float* Ws, float (&acc)[TM / 16][NOUT / 16]) {
  constexpr int RP = TM / 16, NQ = NOUT / 16, NCH = K / KCH, LDW = KCH + 4;
  static_assert(K % KCH == 0 && (NOUT * (KCH / 4)) % DT == 0, "tile shape");
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  stage_weights<NOUT, KCH>(Wg, ldwg, 0, Ws);
  cp_async_commit();
  for (int ch = 0; ch < NCH; ++ch) {
    float* cur = Ws + (ch & 1) * (NOUT * LDW);
    if (ch + 1 < NCH) {
      stage_weights<NOUT, KCH>(Wg, ldwg, (ch + 1) * KCH, Ws + ((ch + 1) & 1) * (NOUT * LDW));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* a_base = As + (ty * RP) * lda + ch * KCH;
#pragma unroll
    for (int kk = 0; kk < KCH; kk += 4) {
      float4 a[RP], b[NQ];
#pragma unroll
      for (int p = 0; p < RP; ++p) a[p] = *reinterpret_cast<const float4*>(a_base + p * lda + kk);
#pragma unroll
      for (int q = 0; q < NQ; ++q) b[q] = *reinterpret_cast<const float4*>(cur + (tx + 16 * q) * LDW + kk);
#pragma unroll
      for (int p = 0; p < RP; ++p)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          acc[p][q] = fmaf(a[p].x, b[q].x, acc[p][q]);
          acc[p][q] = fmaf(a[p].y, b[q].y, acc[p][q]);
          acc[p][q] = fmaf(a[p].z, b[q].z, acc[p][q]);
          acc[p][q] = fmaf(a[p].w, b[q].w, acc[p][q]);
        }
    }
    __syncthreads();
  }
}

template <int RP, int NQ>
__device__ __forceinline__ void zero_acc(float (&acc)[RP][NQ]) {
#pragma unroll
  for (int p = 0; p < RP; ++p)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[p][q] = 0.f;
}

// Coalesced copy of a [rows x W] fp32 tile between global (row stride ldg) and shared (row stride lds).
template <int W>
__device__ __forceinline__ void tile_load(float* S, int lds, const float* __restrict__ G, size_t ldg, int rows) {
  constexpr int V = W / 4;
  for (int id = threadIdx.x; id < rows * V; id += DT) {
    const int r = id / V, c = (id % V) * 4;
    *reinterpret_cast<float4*>(S + r * lds + c) = __ldg(reinterpret_cast<const float4*>(G + (size_t)r * ldg + c));
  }
}
template <int W>
__device__ __forceinline__ void tile_store(float* __restrict__ G, size_t ldg, const float* S, int lds, int rows) {
  constexpr int V = W / 4;
  for (int id = threadIdx.x; id < rows * V; id += DT) {
    const int r = id / V, c = (id % V) * 4;
    *reinterpret_cast<float4*>(G + (size_t)r * ldg + c) = *reinterpret_cast<const float4*>(S + r * lds + c);
  }
}

// Segmented column sums of a [rows x 128] shared tile keyed by a sorted per-row segment id, added
// into out[seg, :].  Runs that touch the first / last row of the tile may continue in a neighbouring
// tile and use atomicAdd (at most two partial sums per segment as long as a segment is shorter than
// the tile, so the result is order-independent); interior runs are plain stores into the
// zero-initialised output.
__device__ __forceinline__ void tile_segment_accumulate(const float* S, int lds, const int* seg, int rows,
                                                        float* __restrict__ out, int width) {
  const int c = threadIdx.x;
  if (c >= width || rows <= 0) return;
  float run = 0.f;
  int cur = seg[0];
  bool first = true;
  for (int r = 0; r < rows; ++r) {
    const int s = seg[r];
    if (s != cur) {
      if (first) atomicAdd(out + (size_t)cur * width + c, run);
      else out[(size_t)cur * width + c] = run;
      first = false;
      run = 0.f;
      cur = s;
    }
    run += S[r * lds + c];
  }
  atomicAdd(out + (size_t)cur * width + c, run);
}

}  // namespace dig3d
