// SphereNet / DimeNet++ dense edge-MLP chain, second generation: TWO 128-edge tiles in flight per SM.
//
// The first-generation chain (spherenet_tc.cu, 3xTF32) is serialisation bound: per 128-edge tile every linear
// costs ~5 k cycles of MMAs and ~6 k cycles of epilogue (TMEM -> bias / swish / residual -> split -> smem),
// strictly alternating, and its fp32-sized operand planes (2 x 66 KB) leave room for one tile per SM only
// (profiles/r01_b_tc_final_ncu_summary.txt: tensor pipe 19.8 % busy).  Here:
//
//   * operands are split into TWO FP16 planes after a power-of-two pre-scale, x*8 = hi + lo (+ 2^-22 x),
//     w*64 = hi + lo: fp16 carries the same 11 significant bits as TF32, so the three-product form
//     D = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi keeps fp32-level accuracy (tools/emulate_split.py: energy error
//     2.2e-7 vs 3.6e-7 for 3xTF32), while kind::f16 MMAs move K = 16 per instruction (twice TF32) and the
//     planes are half the size: two tiles' A operands (132 KB) + a 4-stage weight ring (64 KB) fit in shared
//     memory.  The pre-scales keep `lo` in fp16's normal range for |x| > 0.03 (below that the ABSOLUTE error
//     floor is 4e-9); activations must stay below 8190 in magnitude - a larger value overflows to inf/NaN in
//     the energies and raises the flag read by dig3d_h16_overflow() (the 3xTF32 chain has fp32 range).
//   * the MMA issuer alternates between the two tiles layer by layer: while tile X's eight epilogue warps run
//     the exposed activation phase of layer q, the tensor core runs layer q of tile Y, and vice versa.  The
//     two TMEM accumulators (one K = 64 chunk each, see the truncation note in spherenet_tc.cu) are shared by
//     the tiles; the fp32 residual / skip rows live in the other half of TMEM (2 x 128 columns) instead of
//     registers, which is what lets one epilogue thread own 64 columns of a row.
//   * warp roles (576 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2..9 = epilogue of
//     tile X, warps 10..17 = epilogue of tile Y (two warps per TMEM lane quarter, 64 columns each).
//
// Reference ops: update_e.forward spherenet.py:150-182 (dimenetpp.py:133-161), init.forward spherenet.py:79-91.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc05.cuh"

namespace dig3d {
using namespace tc05;

constexpr int H_M = 128;
constexpr int H_AKU = H_M + 1;                     // k-unit stride of the A planes in 16-byte units (padded)
constexpr int H_PLANE = 16 * H_AKU * 16;           // bytes of one fp16 plane of a [128 x 128] operand tile
constexpr int H_STAGES = 5;                       // 80 KB of weight slabs in flight (the L2 -> smem latency is ~2 k cycles)
constexpr int H_SLAB_K = 32;                       // ring granularity: K = 32 slab = [hi|lo][4 k-units][N][8 halves]
constexpr int H_STAGE_BYTES = 2 * 4 * 128 * 16;    // 16 KB (N = 128)
constexpr int H_TILE_WARPS = 8;
constexpr int H_TILE_THREADS = H_TILE_WARPS * 32;
constexpr int H_CTRL_THREADS = 128;                // warpgroup 0: warp 0 = producer, warp 1 = MMA issuer, warps 2-3 idle
constexpr int H_THREADS = H_CTRL_THREADS + 2 * H_TILE_THREADS;
constexpr int H_CTRL_WARPS = H_CTRL_THREADS / 32;
// Register re-allocation between the warpgroups (setmaxnreg): the kernels are compiled for 96 registers (five warps
// per scheduler); the control warpgroup then drops to 32 and each epilogue warp grows to 112 (the CTA pool is conserved: 128 x 32 + 512 x 112 = 640 x 96; per scheduler 4 x 32 x 112 + 32 x 32
// <= 16384 per scheduler), which is what lets an epilogue thread keep its 64 accumulator values plus the TMEM
// staging registers without spilling (round-2 ncu: the 96-register build spilled the drained accumulators).
__device__ __forceinline__ void h_regs_ctrl() { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); }
__device__ __forceinline__ void h_regs_epi() { asm volatile("setmaxnreg.inc.sync.aligned.u32 112;"); }
constexpr float H_SA = 8.0f, H_SW = 64.0f;
constexpr float H_INV = 1.0f / (H_SA * H_SW);
// |activation| limit of the chain: 65504 / H_SA = 8188
constexpr int H_LDS = 129;                         // row stride (floats) of the fp32 staging overlay of a tile's planes

struct HSmem {
  unsigned char a[2][2][H_PLANE];                  // [tile][hi | lo]
  unsigned char w[H_STAGES][H_STAGE_BYTES];
  float bias[8][128];
  float wr[128 * 8];
  float wr1[64];
  float wr2[128 * 8];                              // fused B + A(next block): lin_rbf2 of the NEXT block (wr is still lin_rbf of this one for the other tile)
  float bias2[2][128];                             //                          b_ji (x 1), b_kj (x H_SA) of the next block
  int dst[2][H_M];
  int aux[2][2][H_M];                              // init_e: atomic numbers of the target / source node of each row
  // d_ready / d_free are indexed [tile][accumulator]: each barrier has ONE waiter that sees every phase in order
  // (a parity wait may never lag or lead its barrier by two phases, which a barrier shared by the tiles allows)
  uint64_t full[H_STAGES], empty[H_STAGES], a_ready[2], d_ready[2][2], d_free[2][2];
  uint64_t l_done;                                 // shared-A mode: every MMA of the layer has completed
  uint32_t tmem_base;
};
static_assert(sizeof(HSmem) <= 227 * 1024, "HSmem exceeds the shared memory of an SM");
static_assert(2 * H_PLANE == H_M * H_LDS * 4, "the fp32 staging overlay must cover exactly one tile's planes");

struct HGemm {
  const unsigned char* w;   // packed slabs (dig3d_h16_pack)
  const float* bias;        // [N] or null
  int K, N;
  int tile_stride;          // bytes added to `w` for the second job of a layer (shared-A mode: the other N half)
};

static __device__ unsigned int g_h16_overflow = 0;
static int h16_fast_swish = 1;
static int h16_wide_epilogue = 0;   // update_e part B (+ A): 0 = eight epilogue warps per tile, 1 = all sixteen on the ready tile
// optional timeline probe: CTA 0 of update_e part B records clock64() at protocol points (tools/gpu_h16_timeline.py)
static __device__ long long g_h16_trace[128];
static __device__ int g_h16_trace_on = 0;
#define H_TRACE(slot) do { if (TRACE && g_h16_trace_on && blockIdx.x == 0) g_h16_trace[(slot)] = clock64(); } while (0)

// x * sigmoid(x).  FAST (default): MUFU ex2 / rcp approximations in their flush-to-zero forms (the non-ftz forms
// cost three extra instructions per element for denormal inputs that cannot occur here); measured on the B200: no
// effect on the energy error (tools/gpu_h16_check.py).  FAST = false: libdevice expf + IEEE division.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FAST>
__device__ __forceinline__ float hswish(float x) {
  return FAST ? x * rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)) : __fdiv_rn(x, 1.0f + expf(-x));
}
// H_SA * swish(t8 / H_SA): the chain keeps its activations pre-scaled by the operand scale, so the split needs no
// multiply and the scale costs nothing (it is folded into the accumulator scale and the bias).
template <bool FAST>
__device__ __forceinline__ float hswish8(float t8) {
  return FAST ? t8 * rcp_ftz(1.0f + ex2_ftz(t8 * (-1.4426950408889634f / H_SA)))
              : __fdiv_rn(t8, 1.0f + expf(-t8 * (1.0f / H_SA)));
}

// ---- producer: every K = 32 slab of every job (layer x tile) through the ring
__device__ __forceinline__ void h_producer_n(HSmem& s, const HGemm* g, int ng, int ntile) {
  int it = 0;
  for (int q = 0; q < ng; ++q) {
    const int nslab = g[q].K / H_SLAB_K;
    const uint32_t bytes = 2u * 4u * (uint32_t)g[q].N * 16u;
    for (int t = 0; t < ntile; ++t)
      for (int c = 0; c < nslab; ++c, ++it) {
        const int st = it % H_STAGES;
        mbar_wait(&s.empty[st], ((it / H_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&s.full[st], bytes);
        bulk_g2s(s.w[st], g[q].w + (size_t)t * g[q].tile_stride + (size_t)c * bytes, bytes, &s.full[st]);
      }
  }
}
template <int NG>
__device__ __forceinline__ void h_producer(HSmem& s, const HGemm (&g)[NG], int ntile) { h_producer_n(s, g, NG, ntile); }

// ---- MMA issuer: jobs alternate between the tiles; one K = 64 chunk (two slabs) per TMEM accumulator.  The two
// accumulators are shared by the tiles: before chunk `ch` overwrites accumulator ch & 1, the tile that used it last
// (chunk ch - 2, possibly the other tile) must have drained it.
template <bool TRACE, bool SHARED_A>
__device__ __forceinline__ void h_mma_n(HSmem& s, const HGemm* g, int ng, int ntile, uint32_t tmem) {
  int it = 0, ch = 0;
  int uses00 = 0, uses01 = 0, uses10 = 0, uses11 = 0;   // uses[tile][accumulator] so far
  int last0 = -1, last1 = -1;                            // tile that used accumulator 0 / 1 last
  for (int q = 0; q < ng; ++q) {
    const int nslab = g[q].K / H_SLAB_K, n = g[q].N;
    const uint32_t idesc = idesc_f16(H_M, n);
    for (int t = 0; t < ntile; ++t) {
      // SHARED_A: the two jobs of a layer are the two N halves of ONE operand tile (K up to 256: the hi plane spans
      // s.a[0], the lo plane s.a[1]), published once per layer by all sixteen epilogue warps
      if (!SHARED_A || t == 0) {
        mbar_wait(&s.a_ready[SHARED_A ? 0 : t], q & 1);
        tc_fence_after();
      }
      if (q < 8) H_TRACE((q * 2 + t) * 2);
      const uint32_t a_hi = smem_u32(SHARED_A ? s.a[0][0] : s.a[t][0]), a_lo = smem_u32(SHARED_A ? s.a[1][0] : s.a[t][1]);
      for (int c = 0; c < nslab; ++c, ++it) {
        const int st = it % H_STAGES, ab = ch & 1;
        if ((c & 1) == 0) {
          const int lt = ab ? last1 : last0;
          if (lt >= 0) {
            const int u = lt ? (ab ? uses11 : uses10) : (ab ? uses01 : uses00);
            mbar_wait(&s.d_free[lt][ab], (u - 1) & 1);   // that tile's epilogue has drained this accumulator
          }
        }
        mbar_wait(&s.full[st], (it / H_STAGES) & 1);
        tc_fence_after();
        const uint32_t w_hi = smem_u32(s.w[st]), w_lo = w_hi + 4u * n * 16u;
        const uint32_t d = tmem + 128u * ab;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {   // the slab's corrections (2^-11 of the main term) first ...
          const uint32_t a_off = (uint32_t)((c * 4 + ks * 2) * H_AKU * 16), b_off = (uint32_t)(ks * 2 * n * 16);
          mma_f16(d, smem_desc(a_lo + a_off, H_AKU * 16, 128), smem_desc(w_hi + b_off, n * 16, 128), idesc,
                  ((c & 1) | ks) != 0);
          mma_f16(d, smem_desc(a_hi + a_off, H_AKU * 16, 128), smem_desc(w_lo + b_off, n * 16, 128), idesc, 1);
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {   // ... then its two hi*hi steps
          const uint32_t a_off = (uint32_t)((c * 4 + ks * 2) * H_AKU * 16), b_off = (uint32_t)(ks * 2 * n * 16);
          mma_f16(d, smem_desc(a_hi + a_off, H_AKU * 16, 128), smem_desc(w_hi + b_off, n * 16, 128), idesc, 1);
        }
        mma_commit(&s.empty[st]);
        if (c & 1) {
          mma_commit(&s.d_ready[t][ab]);
          if (t) { if (ab) ++uses11; else ++uses10; } else { if (ab) ++uses01; else ++uses00; }
          if (ab) last1 = t; else last0 = t;
          ++ch;
        }
      }
      if (q < 8) H_TRACE((q * 2 + t) * 2 + 1);
      if (SHARED_A && t == ntile - 1) mma_commit(&s.l_done);
    }
  }
}
template <int NG, bool TRACE = false>
__device__ __forceinline__ void h_mma(HSmem& s, const HGemm (&g)[NG], int ntile, uint32_t tmem) {
  h_mma_n<TRACE, false>(s, g, NG, ntile, tmem);
}

// ---- epilogue context: thread = one row of its tile, NC = 64 (N = 128) or 32 (N = 64) columns
struct HCtx {
  int t, et, row, half;     // tile in the pair, thread index within the tile's epilogue, row, column half
  uint32_t tl;              // TMEM address of this warp's lane quarter, column 0
  int ntile, ch;            // tiles in this CTA, global accumulator-chunk counter (same sequence as the issuer's)
  int use0, use1;           // chunks of THIS tile drained from accumulator 0 / 1 so far (barrier phases)
  bool bad;                 // a non-finite value reached an OUTPUT of the kernel (fp16 operand range exceeded upstream)
  unsigned char *ahi, *alo; // operand planes this thread writes
};
__device__ __forceinline__ HCtx h_ctx(HSmem& s, int ntile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = (warp - H_CTRL_WARPS) / H_TILE_WARPS, we = (warp - H_CTRL_WARPS) % H_TILE_WARPS;
  const int quarter = warp & 3;   // a warp may only touch TMEM lanes [32*(warp%4), +32)
  return {t, (int)threadIdx.x - H_CTRL_THREADS - t * H_TILE_THREADS, 32 * quarter + lane, we >> 2,
          s.tmem_base + ((uint32_t)(32 * quarter) << 16), ntile, 0, 0, 0, false, s.a[t][0], s.a[t][1]};
}
__device__ __forceinline__ void h_tile_bar(int t) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "n"(H_TILE_THREADS) : "memory");
}
__device__ __forceinline__ void h_epi_done(HSmem& s, int t) {
  fence_async_smem();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&s.a_ready[t]);
}
// 8 consecutive K elements of a row (one k-unit), ALREADY multiplied by H_SA: split into fp16 hi / lo, one 16-byte
// store per plane.  A value beyond the fp16 range becomes inf here and NaN in the products; the kernels flag
// non-finite OUTPUTS (HCtx::bad) instead of paying a range test per operand element.
__device__ __forceinline__ void h_store_ku(HSmem& s, const HCtx& c, int row, int ku, const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  const int o = (ku * H_AKU + row) * 16;
  *reinterpret_cast<uint4*>(c.ahi + o) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(c.alo + o) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void h_store_a16(HSmem& s, const HCtx& c, int col, const float (&v8)[16]) {
  float x[8];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = v8[8 * u + i];
    h_store_ku(s, c, c.row, (col >> 3) + u, x);
  }
}
__device__ __forceinline__ bool h_finite(float x) { return fabsf(x) <= 3.402823466e38f; }
// Cooperative load of a row-major [rows x KU*8] fp32 tile (leading dimension ld floats) into the tile's planes:
// a warp reads 32 consecutive k-units (1 KB when ld == KU*8) per step.
template <int KU>
__device__ __forceinline__ void h_load_tile(HSmem& s, HCtx& c, const float* __restrict__ g, size_t ld, int rows) {
#pragma unroll
  for (int k = 0; k < H_M * KU / H_TILE_THREADS; ++k) {
    const int f = c.et + k * H_TILE_THREADS, row = f / KU, ku = f % KU;
    float x[8];
    if (row < rows) {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * ld + ku * 8));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * ld + ku * 8 + 4));
      x[0] = p0.x * H_SA; x[1] = p0.y * H_SA; x[2] = p0.z * H_SA; x[3] = p0.w * H_SA;
      x[4] = p1.x * H_SA; x[5] = p1.y * H_SA; x[6] = p1.z * H_SA; x[7] = p1.w * H_SA;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = 0.f;
    }
    h_store_ku(s, c, row, ku, x);
  }
}
// The same load in two halves, so that the global-memory round trip of the FIRST operand tile overlaps the kernel's
// set-up (barrier init, TMEM allocation, bias / index loads): fetch before the set-up barrier, commit after it.
// et / t are passed explicitly because the epilogue context does not exist yet.
template <int KU>
struct HTileRegs { float4 v[H_M * KU / H_TILE_THREADS][2]; };
template <int KU>
__device__ __forceinline__ void h_tile_fetch(HTileRegs<KU>& r, int et, const float* __restrict__ g, size_t ld, int rows) {
#pragma unroll
  for (int k = 0; k < H_M * KU / H_TILE_THREADS; ++k) {
    const int f = et + k * H_TILE_THREADS, row = f / KU, ku = f % KU;
    if (row < rows) {
      r.v[k][0] = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * ld + ku * 8));
      r.v[k][1] = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * ld + ku * 8 + 4));
    } else {
      r.v[k][0] = r.v[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
template <int KU>
__device__ __forceinline__ void h_tile_commit(HSmem& s, const HCtx& c, const HTileRegs<KU>& r) {
#pragma unroll
  for (int k = 0; k < H_M * KU / H_TILE_THREADS; ++k) {
    const int f = c.et + k * H_TILE_THREADS, row = f / KU, ku = f % KU;
    const float4 p0 = r.v[k][0], p1 = r.v[k][1];
    const float x[8] = {p0.x * H_SA, p0.y * H_SA, p0.z * H_SA, p0.w * H_SA, p1.x * H_SA, p1.y * H_SA, p1.z * H_SA, p1.w * H_SA};
    h_store_ku(s, c, row, ku, x);
  }
}
// Streaming accumulation of this tile's job: each finished K = 64 chunk is added (round to nearest) into the
// thread's fp32 registers; NP 16-column pieces starting at column col0.  The global chunk counter advances over the
// other tile's chunks of the same layer without touching a barrier (they use that tile's own barriers).
template <int NP, bool FIRST>
__device__ __forceinline__ void h_drain(HSmem& s, HCtx& c, int col0, int chunks, float (&acc)[NP * 16]) {
  if (c.t == 1) c.ch += chunks;
  for (int k = 0; k < chunks; ++k, ++c.ch) {
    const int ab = c.ch & 1;
    const int use = ab ? c.use1 : c.use0;
    mbar_wait(&s.d_ready[c.t][ab], use & 1);
    if (ab) ++c.use1; else ++c.use0;
    tc_fence_after();
    const uint32_t ta = c.tl + 128u * ab + col0;
    if (FIRST && k == 0) {            // straight into the accumulator registers: all loads in flight, one wait
#pragma unroll
      for (int p = 0; p < NP; ++p) tmem_ld16f(ta + 16 * p, &acc[p * 16]);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int p = 0; p < NP; p += 2) {
        uint32_t r[2][16];
        tmem_ld16(ta + 16 * p, r[0]);
        if (p + 1 < NP) tmem_ld16(ta + 16 * p + 16, r[1]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          acc[p * 16 + i] = __fadd_rn(acc[p * 16 + i], __uint_as_float(r[0][i]));
          if (p + 1 < NP) acc[p * 16 + 16 + i] = __fadd_rn(acc[p * 16 + 16 + i], __uint_as_float(r[1][i]));
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&s.d_free[c.t][ab]);   // one arrival per warp
  }
  if (c.t == 0 && c.ntile == 2) c.ch += chunks;
}
__device__ __forceinline__ void h_setup(HSmem& s, int a_ready_warps = H_TILE_WARPS, int d_free_warps = H_TILE_WARPS) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < H_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.a_ready[i], a_ready_warps);
      for (int j = 0; j < 2; ++j) { mbar_init(&s.d_ready[i][j], 1); mbar_init(&s.d_free[i][j], d_free_warps); }
    }
    mbar_init(&s.l_done, 1);
    mbar_fence_init();
  }
  if ((threadIdx.x >> 5) == 0) tmem_alloc(&s.tmem_base, 512);
}
__device__ __forceinline__ void h_finish(HSmem& s, const HCtx* c) {
  if (c && c->bad) atomicOr(&g_h16_overflow, 1u);
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(s.tmem_base, 512);
}
// Row-major [rows x W] fp32 tile (W = 128 or 64) -> global memory with full-line stores: every thread parks its W/2
// values in a [128][W + 1] staging overlay of the tile's planes (lane = row: conflict free), then each warp writes
// whole rows (a direct store from the row-per-thread layout touches 32 lines with 16 bytes each per instruction).
template <int W>
__device__ __forceinline__ void h_store_tile_coalesced(HSmem& s, const HCtx& c, int col0, const float (&v)[W / 2],
                                                       float* __restrict__ out, int rows) {
  constexpr int LD = W + 1, RPW = 128 / W;            // rows written per warp instruction (float4 per lane)
  float* st = reinterpret_cast<float*>(s.a[c.t][0]);
#pragma unroll
  for (int i = 0; i < W / 2; ++i) st[c.row * LD + col0 + i] = v[i];
  h_tile_bar(c.t);
  const int w = c.et >> 5, lane = c.et & 31;
  for (int r0 = w * RPW; r0 < rows; r0 += H_TILE_WARPS * RPW) {
    const int r = r0 + (RPW == 2 ? (lane >> 4) : 0), c4 = 4 * (RPW == 2 ? (lane & 15) : lane);
    if (r < rows) {
      const float* src = st + r * LD + c4;
      *reinterpret_cast<float4*>(out + (size_t)r * W + c4) = make_float4(src[0], src[1], src[2], src[3]);
    }
  }
}

// Same with an output leading dimension (the tile is a column slice of a wider row-major matrix).
// `res` (nullable, same leading dimension): added to the tile in the coalesced store phase (fused residual / skip add).
template <int W>
__device__ __forceinline__ void h_store_tile_strided(HSmem& s, const HCtx& c, int col0, const float (&v)[W / 2],
                                                     float* __restrict__ out, int ld, int rows,
                                                     const float* __restrict__ res = nullptr) {
  constexpr int LD = W + 1, RPW = 128 / W;
  float* st = reinterpret_cast<float*>(s.a[c.t][0]);
#pragma unroll
  for (int i = 0; i < W / 2; ++i) st[c.row * LD + col0 + i] = v[i];
  h_tile_bar(c.t);
  const int w = c.et >> 5, lane = c.et & 31;
  for (int r0 = w * RPW; r0 < rows; r0 += H_TILE_WARPS * RPW) {
    const int r = r0 + (RPW == 2 ? (lane >> 4) : 0), c4 = 4 * (RPW == 2 ? (lane & 15) : lane);
    if (r < rows) {
      const float* src = st + r * LD + c4;
      float4 o = make_float4(src[0], src[1], src[2], src[3]);
      if (res) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(res + (size_t)r * ld + c4));
        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
      }
      *reinterpret_cast<float4*>(out + (size_t)r * ld + c4) = o;
    }
  }
}

// e2 tile (staged as [128][H_LDS] fp32 over the tile's own planes) -> segmented edge -> node sums; the tile's 256
// threads take one column and one half of the rows each.  Rows are target-sorted: only the first and the last
// segment of a half can be shared with another half / tile (atomics), interior segments are plain stores.
__device__ __forceinline__ void h_segment_sums(const HSmem& s, const HCtx& c, int rows, float* __restrict__ v_in) {
  const int col = c.et & 127, r0 = (c.et >> 7) * 64, r1 = min(rows, r0 + 64);
  if (r0 >= r1) return;
  const float* e2t = reinterpret_cast<const float*>(s.a[c.t][0]);
  const int* dst = s.dst[c.t];
  float run = 0.f;
  int cur = dst[r0];
  bool first = true;
  for (int r = r0; r < r1; ++r) {
    const int d = dst[r];
    if (d != cur) {
      if (first) atomicAdd(v_in + (size_t)cur * 128 + col, run);
      else v_in[(size_t)cur * 128 + col] = run;
      first = false; run = 0.f; cur = d;
    }
    run += e2t[r * H_LDS + col];
  }
  atomicAdd(v_in + (size_t)cur * 128 + col, run);
}

// ---------------------------------------------------------------------------------- weight packing
// W [N, K] fp32 (nn.Linear layout) -> K/32 slabs of [hi|lo][4 k-units][N][8 halves], w*64 = hi + lo.
struct HPackJob { const float* w; unsigned char* out; int N, K, trans; };   // trans > 0: the source is stored [K, trans] (a column slice of W^T's source)
struct HPackJobs { HPackJob job[16]; };
__global__ void h16_pack_kernel(HPackJobs jobs) {
  const HPackJob jb = jobs.job[blockIdx.y];
  const int total = jb.N * jb.K;
  for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
    const int n = id / jb.K, k = id % jb.K;
    const float x = __ldg(jb.w + (jb.trans ? (size_t)k * jb.trans + n : (size_t)id)) * H_SW;
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    const int c = k >> 5, ku = (k & 31) >> 3, kk = k & 7;
    const size_t slab = (size_t)c * (2 * 4 * jb.N * 16);
    const size_t o = ((size_t)ku * jb.N + n) * 16 + kk * 2;
    *reinterpret_cast<__half*>(jb.out + slab + o) = h;
    *reinterpret_cast<__half*>(jb.out + slab + (size_t)4 * jb.N * 16 + o) = l;
  }
}

// ---------------------------------------------------------------------------------- update_e part B
struct HBParams {
  HGemm g[11];                 // lin_up, res0.lin1, res0.lin2, lin, res1.lin1, res1.lin2, res2.lin1, res2.lin2
                               // FUSE: + lin_ji, lin_kj, lin_down of the NEXT block
  const float* w_rbf;          // [128, 6]
  const float *n_w_rbf1, *n_w_rbf2;     // FUSE: lin_rbf1 [8, 6], lin_rbf2 [128, 8] of the next block
  float *n_x_ji, *n_x_down;             // FUSE: part-A outputs of the next block
};

// FUSE: part A of the NEXT interaction block (x_ji = act(lin_ji(e1)), x_kj = act(lin_kj(e1)) * gate,
// x_down = act(lin_down(x_kj)); spherenet.py:154-161) is appended to this block's chain: its operand is the e1 tile
// this kernel has just produced, so the three extra jobs continue on the same tiles -- one launch, one set-up and
// one e1 read less per block (part A alone is 34 us for 2.5 layers, two thirds of it set-up and tail).
template <bool FAST, bool FUSE>
__global__ void __launch_bounds__(H_THREADS, 1)
sphere_update_e_b_h16_kernel(const float* __restrict__ m, const float* __restrict__ x_ji,
                             const float* __restrict__ e1_in, const float* __restrict__ rbf0,
                             const int32_t* __restrict__ dst, int n_edges, HBParams P, float* __restrict__ e1_out,
                             float* __restrict__ v_in) {
  extern __shared__ __align__(1024) unsigned char h_raw[];
  HSmem& s = *reinterpret_cast<HSmem*>(h_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_edges + H_M - 1) / H_M, tile0 = blockIdx.x * 2, ntile = min(2, n_tiles - tile0);
  if (tid == 0 && g_h16_trace_on && blockIdx.x == 0) { g_h16_trace[100] = clock64(); g_h16_trace[101] = (long long)global_ns(); }
  HTileRegs<8> m_regs;                              // the m tile (K = 64) of this thread's tile, in flight during set-up
  if (warp >= H_CTRL_WARPS) {
    const int t = (warp - H_CTRL_WARPS) / H_TILE_WARPS, et = tid - H_CTRL_THREADS - t * H_TILE_THREADS;
    const int e0 = (tile0 + t) * H_M;
    if (t < ntile) h_tile_fetch<8>(m_regs, et, m + (size_t)e0 * 64, 64, min(H_M, n_edges - e0));
  }
  h_setup(s);
  for (int i = tid; i < 8 * 128; i += H_THREADS) {
    const float* b = P.g[i / 128].bias;
    s.bias[i / 128][i % 128] = b ? H_SA * __ldg(b + i % 128) : 0.f;          // the chain runs pre-scaled by H_SA
  }
  for (int i = tid; i < 128 * 8; i += H_THREADS) s.wr[i] = (i % 8 < 6) ? __ldg(P.w_rbf + (i / 8) * 6 + i % 8) : 0.f;
  for (int i = tid; i < 2 * H_M; i += H_THREADS) {
    const int e = (tile0 + i / H_M) * H_M + i % H_M;
    s.dst[i / H_M][i % H_M] = (e < n_edges) ? dst[e] : -1;
  }
  if (FUSE) {
    for (int i = tid; i < 2 * 128; i += H_THREADS)     // lin_ji's output leaves unscaled, lin_kj's feeds an operand (x H_SA)
      s.bias2[i / 128][i % 128] = (i / 128 ? H_SA : 1.0f) * __ldg(P.g[8 + i / 128].bias + i % 128);
    for (int i = tid; i < 128 * 8; i += H_THREADS) s.wr2[i] = __ldg(P.n_w_rbf2 + i);      // [128][8]
    for (int i = tid; i < 64; i += H_THREADS) s.wr1[i] = (i % 8 < 6) ? __ldg(P.n_w_rbf1 + (i / 8) * 6 + i % 8) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0 && g_h16_trace_on && blockIdx.x == 0) g_h16_trace[104] = clock64();
  HCtx c;
  bool epi = false;
  constexpr int NG = FUSE ? 11 : 8;
  if (warp < H_CTRL_WARPS) {
    h_regs_ctrl();
    if (tid == 0) h_producer_n(s, P.g, NG, ntile);
    else if (tid == 32) h_mma_n<true, false>(s, P.g, NG, ntile, s.tmem_base);
  } else if (h_regs_epi(), (c = h_ctx(s, ntile)).t < ntile) {
    epi = true;
    constexpr bool TRACE = true;
    const bool probe = c.et == 0;
    if (probe && c.t == 0) H_TRACE(105);
    const int e0 = (tile0 + c.t) * H_M, rows = min(H_M, n_edges - e0);
    const bool valid = c.row < rows;
    const size_t ge = (size_t)(e0 + c.row);
    const int col0 = c.half * 64;
    const uint32_t stash = c.tl + 256u + 128u * c.t + col0;
    // The fp32 skip rows (x_ji for q = 0, e1_in for q = 3) are fetched into the TMEM stash while this thread would
    // otherwise wait for its tile's MMAs, so no global load sits on the critical path of an epilogue.
    auto prefetch_skip = [&](const float* __restrict__ src) {
      const float* gsrc = src + ge * 128 + col0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 x = valid ? __ldg(reinterpret_cast<const float4*>(gsrc + 16 * p + i)) : make_float4(0, 0, 0, 0);
          r[i] = __float_as_uint(x.x * H_SA); r[i + 1] = __float_as_uint(x.y * H_SA);
          r[i + 2] = __float_as_uint(x.z * H_SA); r[i + 3] = __float_as_uint(x.w * H_SA);
        }
        tmem_st16(stash + 16 * p, r);
      }
      tmem_st_wait();
    };
    // A0 = m tile (K = 64), fetched before the set-up barrier
    h_tile_commit<8>(s, c, m_regs);
    if (probe && c.t == 0) H_TRACE(106);
    h_epi_done(s, c.t);
    if (probe && c.t == 0) H_TRACE(107);
    prefetch_skip(x_ji);
    if (probe && c.t == 0) H_TRACE(108);
    // The eight epilogues of the chain (spherenet.py:172-179); stash = the fp32 skip / residual row (x H_SA):
    //   q=0: h = stash(x_ji) + act(lin_up(m))                -> A, stash
    //   q=1,4,6: t = act(lin1(h))                            -> A
    //   q=2,5: h = stash + act(lin2(t))                      -> A, stash      (q=2: then stash <- e1_in)
    //   q=3: h = act(lin(h)) + stash(e1_in)                  -> A, stash
    //   q=7: h = stash + act(lin2(t))                        -> e1_out, e2 tile
    float acc_final[64];
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
      float (&acc)[64] = acc_final;
      if (probe) H_TRACE(32 + c.t * 32 + q * 4);
      h_drain<4, true>(s, c, col0, q == 0 ? 1 : 2, acc);
      if (probe) H_TRACE(32 + c.t * 32 + q * 4 + 1);
      const bool add_stash = (q == 0 || q == 2 || q == 3 || q == 5 || q == 7);
      const bool to_stash = (q == 0 || q == 3 || q == 5);
      float rb[6];
      if (q == 7) {
#pragma unroll
        for (int n = 0; n < 6; ++n) rb[n] = valid ? __ldg(rbf0 + ge * 6 + n) : 0.f;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int col = col0 + 16 * p;
        uint32_t r[16];
        if (add_stash) tmem_ld16(stash + 16 * p, r);           // in flight under the 16 activations below
        float* v = &acc[16 * p];                               // in place: v8 = H_SA * act(.)
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&s.bias[q][col + i]);
          v[i] = hswish8<FAST>(fmaf(v[i], H_SA * H_INV, b.x));
          v[i + 1] = hswish8<FAST>(fmaf(v[i + 1], H_SA * H_INV, b.y));
          v[i + 2] = hswish8<FAST>(fmaf(v[i + 2], H_SA * H_INV, b.z));
          v[i + 3] = hswish8<FAST>(fmaf(v[i + 3], H_SA * H_INV, b.w));
        }
        if (add_stash) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(r[i]);
        }
        if (q < 7) {
          if (to_stash) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
            tmem_st16(stash + 16 * p, r);
          }
          float v16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v16[i] = v[i];
          h_store_a16(s, c, col, v16);
        } else {
          // e1 (unscaled, kept in the accumulator registers for the coalesced store below) and
          // e2 = lin_rbf(rbf0) * e1, staged as a [128][H_LDS] tile over this tile's planes (all its MMAs are done)
          float* e2t = reinterpret_cast<float*>(s.a[c.t][0]);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float o = v[i] * (1.0f / H_SA);
            c.bad |= !h_finite(o);
            const float4 w0 = *reinterpret_cast<const float4*>(s.wr + (col + i) * 8);
            const float2 w1 = *reinterpret_cast<const float2*>(s.wr + (col + i) * 8 + 4);
            const float gsum = fmaf(w1.y, rb[5], fmaf(w1.x, rb[4], fmaf(w0.w, rb[3], fmaf(w0.z, rb[2],
                               fmaf(w0.y, rb[1], w0.x * rb[0])))));
            e2t[c.row * H_LDS + col + i] = gsum * o;
            v[i] = o;
          }
        }
      }
      if (probe) H_TRACE(32 + c.t * 32 + q * 4 + 2);
      if (q < 7) {
        if (to_stash) tmem_st_wait();
        h_epi_done(s, c.t);
        if (q == 2) prefetch_skip(e1_in);     // the stash was consumed above; q = 3 adds e1_in from it
      }
      if (probe) H_TRACE(32 + c.t * 32 + q * 4 + 3);
    }
    h_tile_bar(c.t);
    h_segment_sums(s, c, rows, v_in);   // spherenet.py:211
    h_tile_bar(c.t);
    h_store_tile_coalesced<128>(s, c, col0, acc_final, e1_out + (size_t)e0 * 128, rows);
    if (probe) H_TRACE(96 + c.t);
    if (FUSE) {
      // ---- part A of the next block on the e1 tile still in this thread's registers
      h_tile_bar(c.t);                               // every thread of the tile has read the staging overlay
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float v16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v16[i] = acc_final[16 * p + i] * H_SA;
        h_store_a16(s, c, col0 + 16 * p, v16);
      }
      float r8[8];                                   // gate coefficients of this row: lin_rbf1(rbf0[row])   spherenet.py:157
      {
        float rb[6];
#pragma unroll
        for (int n = 0; n < 6; ++n) rb[n] = valid ? __ldg(rbf0 + ge * 6 + n) : 0.f;
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
          float a = 0.f;
#pragma unroll
          for (int n = 0; n < 6; ++n) a = fmaf(s.wr1[mm * 8 + n], rb[n], a);
          r8[mm] = a;
        }
      }
      h_epi_done(s, c.t);
      float (&acc)[64] = acc_final;
      // G0: x_ji = act(lin_ji(e1))                                                spherenet.py:154
      h_drain<4, true>(s, c, col0, 2, acc);
      h_epi_done(s, c.t);   // the operand (= e1) is reused unchanged by lin_kj
      if (valid) {
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&s.bias2[0][col0 + i]);
          float4 o;
          o.x = hswish<FAST>(fmaf(acc[i], H_INV, b.x));
          o.y = hswish<FAST>(fmaf(acc[i + 1], H_INV, b.y));
          o.z = hswish<FAST>(fmaf(acc[i + 2], H_INV, b.z));
          o.w = hswish<FAST>(fmaf(acc[i + 3], H_INV, b.w));
          *reinterpret_cast<float4*>(P.n_x_ji + ge * 128 + col0 + i) = o;
        }
      }
      // G1: x_kj = act(lin_kj(e1)) * lin_rbf2(r8)                                 spherenet.py:155-159
      h_drain<4, true>(s, c, col0, 2, acc);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = col0 + 16 * p + i;
          const float4 w0 = *reinterpret_cast<const float4*>(s.wr2 + col * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(s.wr2 + col * 8 + 4);
          const float gate = fmaf(w1.w, r8[7], fmaf(w1.z, r8[6], fmaf(w1.y, r8[5], fmaf(w1.x, r8[4],
                             fmaf(w0.w, r8[3], fmaf(w0.z, r8[2], fmaf(w0.y, r8[1], w0.x * r8[0])))))));
          v[i] = hswish8<FAST>(fmaf(acc[16 * p + i], H_SA * H_INV, s.bias2[1][col])) * gate;
        }
        h_store_a16(s, c, col0 + 16 * p, v);
      }
      h_epi_done(s, c.t);
      // G2: x_down = act(lin_down(x_kj)), N = 64                                  spherenet.py:161
      {
        const int col = c.half * 32;
        float a32[32];
        h_drain<2, true>(s, c, col, 2, a32);
#pragma unroll
        for (int i = 0; i < 32; ++i) a32[i] = hswish<FAST>(a32[i] * H_INV);
        h_store_tile_coalesced<64>(s, c, col, a32, P.n_x_down + (size_t)e0 * 64, rows);   // all MMAs of the tile are done
      }
    }
  }
  h_finish(s, epi ? &c : nullptr);
  if (tid == 0 && g_h16_trace_on && blockIdx.x == 0) { g_h16_trace[102] = clock64(); g_h16_trace[103] = (long long)global_ns(); }
}

// ---------------------------------------------------------------------------------- update_e part B (+ A), wide epilogue
// Same chain, same jobs, same barriers as sphere_update_e_b_h16_kernel; the difference is WHO runs an epilogue.  There,
// eight warps own a tile (64 columns per thread) and a tile's job cycle is its MMAs (~3.6 k cycles) PLUS its epilogue
// (~5 k: 2 k of MUFU and 1.8 k of FP32 pipe per scheduler on two warps) -- the other tile fills the gaps and the tensor
// pipe still idles a quarter of the time.  Here all SIXTEEN epilogue warps serve the tile whose
// accumulators are ready (32 columns per thread, four warps per scheduler), then the other tile: an epilogue takes about
// half as long, MMA(X, q+1) can follow MMA(Y, q) immediately, and the cycle becomes MMA-issue bound.  A thread walks the
// jobs in the issuer's order (q, tile 0), (q, tile 1), (q + 1, tile 0), ...; nothing but barrier phases is kept per tile.
constexpr int X_WARPS = 2 * H_TILE_WARPS;         // epilogue warps, all on one tile at a time
constexpr int X_THREADS = X_WARPS * 32;

struct XCtx {
  int e, row, slice;        // index among the 512 epilogue threads, row of the tile (= TMEM lane), 32-column slice
  uint32_t tl;              // TMEM address of this warp's lane quarter, column 0
  int ch;                   // accumulator-chunk counter (the issuer's sequence: q major, tile minor)
  int use[2][2];            // chunks drained so far per [tile][accumulator]: barrier phases
  bool bad;
};
__device__ __forceinline__ void x_bar() { asm volatile("bar.sync 1, %0;" ::"n"(X_THREADS) : "memory"); }
__device__ __forceinline__ void x_epi_done(HSmem& s, int t) {
  fence_async_smem();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&s.a_ready[t]);
}
// one k-unit (8 consecutive K elements, already x H_SA) of row `row` of tile t -> fp16 hi / lo planes
__device__ __forceinline__ void x_store_ku(HSmem& s, int t, int row, int ku, const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  const int o = (ku * H_AKU + row) * 16;
  *reinterpret_cast<uint4*>(s.a[t][0] + o) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(s.a[t][1] + o) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void x_store_a16(HSmem& s, const XCtx& c, int t, int col, const float (&v8)[16]) {
  float x[8];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = v8[8 * u + i];
    x_store_ku(s, t, c.row, (col >> 3) + u, x);
  }
}
// chunks of job (q, t) -> fp32 registers (round-to-nearest adds), NP 16-column pieces from column col0
template <int NP>
__device__ __forceinline__ void x_drain(HSmem& s, XCtx& c, int t, int col0, int chunks, float (&acc)[NP * 16]) {
  for (int k = 0; k < chunks; ++k, ++c.ch) {
    const int ab = c.ch & 1;
    mbar_wait(&s.d_ready[t][ab], c.use[t][ab] & 1);
    ++c.use[t][ab];
    tc_fence_after();
    const uint32_t ta = c.tl + 128u * ab + col0;
    if (k == 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p) tmem_ld16f(ta + 16 * p, &acc[p * 16]);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        uint32_t r[16];
        tmem_ld16(ta + 16 * p, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[p * 16 + i] = __fadd_rn(acc[p * 16 + i], __uint_as_float(r[i]));
      }
    }
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&s.d_free[t][ab]);
  }
}
// [rows x W] fp32 tile of tile t (every thread holds W / 4 values of its row from column col0) -> full-line stores
template <int W>
__device__ __forceinline__ void x_store_tile_coalesced(HSmem& s, const XCtx& c, int t, int col0, const float (&v)[W / 4],
                                                       float* __restrict__ out, int rows) {
  constexpr int LD = W + 1, RPW = 128 / W;
  float* st = reinterpret_cast<float*>(s.a[t][0]);
#pragma unroll
  for (int i = 0; i < W / 4; ++i) st[c.row * LD + col0 + i] = v[i];
  x_bar();
  const int w = c.e >> 5, lane = c.e & 31;
  for (int r0 = w * RPW; r0 < rows; r0 += X_WARPS * RPW) {
    const int r = r0 + (RPW == 2 ? (lane >> 4) : 0), c4 = 4 * (RPW == 2 ? (lane & 15) : lane);
    if (r < rows) {
      const float* src = st + r * LD + c4;
      *reinterpret_cast<float4*>(out + (size_t)r * W + c4) = make_float4(src[0], src[1], src[2], src[3]);
    }
  }
}
// e2 tile of tile t (staged [128][H_LDS] over its planes) -> edge -> node sums: one column and one 32-row quarter per thread
__device__ __forceinline__ void x_segment_sums(const HSmem& s, const XCtx& c, int t, int rows, float* __restrict__ v_in) {
  const int col = c.e & 127, r0 = (c.e >> 7) * 32, r1 = min(rows, r0 + 32);
  if (r0 >= r1) return;
  const float* e2t = reinterpret_cast<const float*>(s.a[t][0]);
  const int* dst = s.dst[t];
  float run = 0.f;
  int cur = dst[r0];
  bool first = true;
  for (int r = r0; r < r1; ++r) {
    const int d = dst[r];
    if (d != cur) {
      if (first) atomicAdd(v_in + (size_t)cur * 128 + col, run);
      else v_in[(size_t)cur * 128 + col] = run;
      first = false; run = 0.f; cur = d;
    }
    run += e2t[r * H_LDS + col];
  }
  atomicAdd(v_in + (size_t)cur * 128 + col, run);
}

template <bool FAST, bool FUSE>
__global__ void __launch_bounds__(H_THREADS, 1)
sphere_update_e_b_x16_kernel(const float* __restrict__ m, const float* __restrict__ x_ji,
                             const float* __restrict__ e1_in, const float* __restrict__ rbf0,
                             const int32_t* __restrict__ dst, int n_edges, HBParams P, float* __restrict__ e1_out,
                             float* __restrict__ v_in) {
  extern __shared__ __align__(1024) unsigned char h_raw[];
  HSmem& s = *reinterpret_cast<HSmem*>(h_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_edges + H_M - 1) / H_M, tile0 = blockIdx.x * 2, ntile = min(2, n_tiles - tile0);
  // the m tiles (K = 64: 8 k-units per row) of both tiles, in flight during the set-up: 2 (row, k-unit) items per tile
  float4 mreg[2][2][2];
  if (warp >= H_CTRL_WARPS) {
    const int e = tid - H_CTRL_THREADS;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int f = e + k * X_THREADS, row = f >> 3, ku = f & 7;
        const int e0 = (tile0 + t) * H_M, rows = min(H_M, n_edges - e0);
        if (t < ntile && row < rows) {
          const float* g = m + (size_t)(e0 + row) * 64 + ku * 8;
          mreg[t][k][0] = __ldg(reinterpret_cast<const float4*>(g));
          mreg[t][k][1] = __ldg(reinterpret_cast<const float4*>(g + 4));
        } else {
          mreg[t][k][0] = mreg[t][k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
  }
  h_setup(s, X_WARPS, X_WARPS);
  for (int i = tid; i < 8 * 128; i += H_THREADS) {
    const float* b = P.g[i / 128].bias;
    s.bias[i / 128][i % 128] = b ? H_SA * __ldg(b + i % 128) : 0.f;          // the chain runs pre-scaled by H_SA
  }
  for (int i = tid; i < 128 * 8; i += H_THREADS) s.wr[i] = (i % 8 < 6) ? __ldg(P.w_rbf + (i / 8) * 6 + i % 8) : 0.f;
  for (int i = tid; i < 2 * H_M; i += H_THREADS) {
    const int e = (tile0 + i / H_M) * H_M + i % H_M;
    s.dst[i / H_M][i % H_M] = (e < n_edges) ? dst[e] : -1;
  }
  if (FUSE) {
    for (int i = tid; i < 2 * 128; i += H_THREADS)     // lin_ji's output leaves unscaled, lin_kj's feeds an operand (x H_SA)
      s.bias2[i / 128][i % 128] = (i / 128 ? H_SA : 1.0f) * __ldg(P.g[8 + i / 128].bias + i % 128);
    for (int i = tid; i < 128 * 8; i += H_THREADS) s.wr2[i] = __ldg(P.n_w_rbf2 + i);      // [128][8]
    for (int i = tid; i < 64; i += H_THREADS) s.wr1[i] = (i % 8 < 6) ? __ldg(P.n_w_rbf1 + (i / 8) * 6 + i % 8) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr int NG = FUSE ? 11 : 8;
  XCtx c;
  bool epi = false;
  if (warp < H_CTRL_WARPS) {
    h_regs_ctrl();
    if (tid == 0) h_producer_n(s, P.g, NG, ntile);
    else if (tid == 32) h_mma_n<false, false>(s, P.g, NG, ntile, s.tmem_base);
  } else {
    h_regs_epi();
    epi = true;
    const int lane = tid & 31, ew = warp - H_CTRL_WARPS, quarter = warp & 3;
    c.e = tid - H_CTRL_THREADS; c.row = 32 * quarter + lane; c.slice = ew >> 2;
    c.tl = s.tmem_base + ((uint32_t)(32 * quarter) << 16);
    c.ch = 0; c.use[0][0] = c.use[0][1] = c.use[1][0] = c.use[1][1] = 0; c.bad = false;
    const int col0 = 32 * c.slice;
    // skip rows (x_ji for q = 0, e1_in for q = 3) -> TMEM stash of tile t, x H_SA
    auto prefetch_skip = [&](const float* __restrict__ src, int t) {
      const int e0 = (tile0 + t) * H_M;
      const bool valid = c.row < min(H_M, n_edges - e0);
      const float* gsrc = src + (size_t)(e0 + c.row) * 128 + col0;
      const uint32_t stash = c.tl + 256u + 128u * t + col0;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 x = valid ? __ldg(reinterpret_cast<const float4*>(gsrc + 16 * p + i)) : make_float4(0, 0, 0, 0);
          r[i] = __float_as_uint(x.x * H_SA); r[i + 1] = __float_as_uint(x.y * H_SA);
          r[i + 2] = __float_as_uint(x.z * H_SA); r[i + 3] = __float_as_uint(x.w * H_SA);
        }
        tmem_st16(stash + 16 * p, r);
      }
      tmem_st_wait();
    };
    // A0 = the m tiles
    for (int t = 0; t < ntile; ++t) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int f = c.e + k * X_THREADS, row = f >> 3, ku = f & 7;
        const float4 p0 = mreg[t][k][0], p1 = mreg[t][k][1];
        const float x[8] = {p0.x * H_SA, p0.y * H_SA, p0.z * H_SA, p0.w * H_SA, p1.x * H_SA, p1.y * H_SA, p1.z * H_SA, p1.w * H_SA};
        x_store_ku(s, t, row, ku, x);
      }
      x_epi_done(s, t);
    }
    for (int t = 0; t < ntile; ++t) prefetch_skip(x_ji, t);
    float acc[32];
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
      const bool add_stash = (q == 0 || q == 2 || q == 3 || q == 5 || q == 7);
      const bool to_stash = (q == 0 || q == 3 || q == 5);
#pragma unroll 1
      for (int t = 0; t < ntile; ++t) {
        const int e0 = (tile0 + t) * H_M, rows = min(H_M, n_edges - e0);
        const bool valid = c.row < rows;
        const size_t ge = (size_t)(e0 + c.row);
        const uint32_t stash = c.tl + 256u + 128u * t + col0;
        x_drain<2>(s, c, t, col0, q == 0 ? 1 : 2, acc);
        float rb[6];
        if (q == 7) {
#pragma unroll
          for (int n = 0; n < 6; ++n) rb[n] = valid ? __ldg(rbf0 + ge * 6 + n) : 0.f;
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int col = col0 + 16 * p;
          uint32_t r[16];
          if (add_stash) tmem_ld16(stash + 16 * p, r);           // in flight under the 16 activations below
          float* v = &acc[16 * p];                               // in place: v8 = H_SA * act(.)
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(&s.bias[q][col + i]);
            v[i] = hswish8<FAST>(fmaf(v[i], H_SA * H_INV, b.x));
            v[i + 1] = hswish8<FAST>(fmaf(v[i + 1], H_SA * H_INV, b.y));
            v[i + 2] = hswish8<FAST>(fmaf(v[i + 2], H_SA * H_INV, b.z));
            v[i + 3] = hswish8<FAST>(fmaf(v[i + 3], H_SA * H_INV, b.w));
          }
          if (add_stash) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(r[i]);
          }
          if (q < 7) {
            if (to_stash) {
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
              tmem_st16(stash + 16 * p, r);
            }
            float v16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v16[i] = v[i];
            x_store_a16(s, c, t, col, v16);
          } else {
            // e1 (unscaled, kept in the accumulator registers for the coalesced store below) and e2 = lin_rbf(rbf0) * e1,
            // staged as a [128][H_LDS] tile over this tile's planes (all its MMAs of part B are done)
            float* e2t = reinterpret_cast<float*>(s.a[t][0]);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float o = v[i] * (1.0f / H_SA);
              c.bad |= !h_finite(o);
              const float4 w0 = *reinterpret_cast<const float4*>(s.wr + (col + i) * 8);
              const float2 w1 = *reinterpret_cast<const float2*>(s.wr + (col + i) * 8 + 4);
              const float gsum = fmaf(w1.y, rb[5], fmaf(w1.x, rb[4], fmaf(w0.w, rb[3], fmaf(w0.z, rb[2],
                                 fmaf(w0.y, rb[1], w0.x * rb[0])))));
              e2t[c.row * H_LDS + col + i] = gsum * o;
              v[i] = o;
            }
          }
        }
        if (q < 7) {
          if (to_stash) tmem_st_wait();
          x_epi_done(s, t);
          if (q == 2) prefetch_skip(e1_in, t);    // the stash was consumed above; q = 3 adds e1_in from it
        } else {
          x_bar();
          x_segment_sums(s, c, t, rows, v_in);    // spherenet.py:211
          x_bar();
          x_store_tile_coalesced<128>(s, c, t, col0, acc, e1_out + (size_t)e0 * 128, rows);
          if (FUSE) {
            // operand of part A of the next block: H_SA * e1, from the registers
            x_bar();                              // every thread has read the staging overlay
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              float v16[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] = acc[16 * p + i] * H_SA;
              x_store_a16(s, c, t, col0 + 16 * p, v16);
            }
            x_epi_done(s, t);
          }
        }
      }
    }
    if (FUSE) {
      // ---- part A of the next block (spherenet.py:154-161): three more jobs per tile
      // G0: x_ji = act(lin_ji(e1))
#pragma unroll 1
      for (int t = 0; t < ntile; ++t) {
        const int e0 = (tile0 + t) * H_M, rows = min(H_M, n_edges - e0);
        const bool valid = c.row < rows;
        const size_t ge = (size_t)(e0 + c.row);
        x_drain<2>(s, c, t, col0, 2, acc);
        x_epi_done(s, t);     // the operand (= e1) is reused unchanged by lin_kj
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(&s.bias2[0][col0 + i]);
            float4 o;
            o.x = hswish<FAST>(fmaf(acc[i], H_INV, b.x));
            o.y = hswish<FAST>(fmaf(acc[i + 1], H_INV, b.y));
            o.z = hswish<FAST>(fmaf(acc[i + 2], H_INV, b.z));
            o.w = hswish<FAST>(fmaf(acc[i + 3], H_INV, b.w));
            *reinterpret_cast<float4*>(P.n_x_ji + ge * 128 + col0 + i) = o;
          }
        }
      }
      // G1: x_kj = act(lin_kj(e1)) * lin_rbf2(lin_rbf1(rbf0))
#pragma unroll 1
      for (int t = 0; t < ntile; ++t) {
        const int e0 = (tile0 + t) * H_M, rows = min(H_M, n_edges - e0);
        const bool valid = c.row < rows;
        const size_t ge = (size_t)(e0 + c.row);
        float r8[8];
        {
          float rb[6];
#pragma unroll
          for (int n = 0; n < 6; ++n) rb[n] = valid ? __ldg(rbf0 + ge * 6 + n) : 0.f;
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) {
            float a = 0.f;
#pragma unroll
            for (int n = 0; n < 6; ++n) a = fmaf(s.wr1[mm * 8 + n], rb[n], a);
            r8[mm] = a;
          }
        }
        x_drain<2>(s, c, t, col0, 2, acc);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = col0 + 16 * p + i;
            const float4 w0 = *reinterpret_cast<const float4*>(s.wr2 + col * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(s.wr2 + col * 8 + 4);
            const float gate = fmaf(w1.w, r8[7], fmaf(w1.z, r8[6], fmaf(w1.y, r8[5], fmaf(w1.x, r8[4],
                               fmaf(w0.w, r8[3], fmaf(w0.z, r8[2], fmaf(w0.y, r8[1], w0.x * r8[0])))))));
            v[i] = hswish8<FAST>(fmaf(acc[16 * p + i], H_SA * H_INV, s.bias2[1][col])) * gate;
          }
          x_store_a16(s, c, t, col0 + 16 * p, v);
        }
        x_epi_done(s, t);
      }
      // G2: x_down = act(lin_down(x_kj)), N = 64: 16 columns per thread
#pragma unroll 1
      for (int t = 0; t < ntile; ++t) {
        const int e0 = (tile0 + t) * H_M, rows = min(H_M, n_edges - e0);
        const int col = 16 * c.slice;
        float a16[16];
        x_drain<1>(s, c, t, col, 2, a16);
#pragma unroll
        for (int i = 0; i < 16; ++i) a16[i] = hswish<FAST>(a16[i] * H_INV);
        x_store_tile_coalesced<64>(s, c, t, col, a16, P.n_x_down + (size_t)e0 * 64, rows);   // all MMAs of the tile are done
        x_bar();              // the staging overlay of tile t is read before anything else may touch shared memory
      }
    }
    if (c.bad) atomicOr(&g_h16_overflow, 1u);
  }
  (void)epi;
  h_finish(s, nullptr);
}

// ---------------------------------------------------------------------------------- update_e part A
struct HAParams {
  HGemm g[3];                  // lin_ji, lin_kj, lin_down
  const float *w_rbf1, *w_rbf2;
};

template <bool FAST>
__global__ void __launch_bounds__(H_THREADS, 1)
sphere_update_e_a_h16_kernel(const float* __restrict__ e1, const float* __restrict__ rbf0, int n_edges, HAParams P,
                             float* __restrict__ x_ji, float* __restrict__ x_down) {
  extern __shared__ __align__(1024) unsigned char h_raw[];
  HSmem& s = *reinterpret_cast<HSmem*>(h_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_edges + H_M - 1) / H_M, tile0 = blockIdx.x * 2, ntile = min(2, n_tiles - tile0);
  h_setup(s);
  for (int i = tid; i < 2 * 128; i += H_THREADS)     // lin_ji's output leaves unscaled, lin_kj's feeds an operand (x H_SA)
    s.bias[i / 128][i % 128] = (i / 128 ? H_SA : 1.0f) * __ldg(P.g[i / 128].bias + i % 128);
  for (int i = tid; i < 128 * 8; i += H_THREADS) s.wr[i] = __ldg(P.w_rbf2 + i);      // [128][8]
  for (int i = tid; i < 64; i += H_THREADS) s.wr1[i] = (i % 8 < 6) ? __ldg(P.w_rbf1 + (i / 8) * 6 + i % 8) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  HCtx c;
  bool epi = false;
  if (warp < H_CTRL_WARPS) {
    h_regs_ctrl();
    if (tid == 0) h_producer(s, P.g, ntile);
    else if (tid == 32) h_mma(s, P.g, ntile, s.tmem_base);
  } else if (h_regs_epi(), (c = h_ctx(s, ntile)).t < ntile) {
    epi = true;
    const int e0 = (tile0 + c.t) * H_M, rows = min(H_M, n_edges - e0);
    const bool valid = c.row < rows;
    const size_t ge = (size_t)(e0 + c.row);
    const int col0 = c.half * 64;
    h_load_tile<16>(s, c, e1 + (size_t)e0 * 128, 128, rows);
    // rbf gate coefficients of this row: r8 = lin_rbf1(rbf0[row])                spherenet.py:157
    float r8[8];
    {
      float rb[6];
#pragma unroll
      for (int n = 0; n < 6; ++n) rb[n] = valid ? __ldg(rbf0 + ge * 6 + n) : 0.f;
#pragma unroll
      for (int mm = 0; mm < 8; ++mm) {
        float a = 0.f;
#pragma unroll
        for (int n = 0; n < 6; ++n) a = fmaf(s.wr1[mm * 8 + n], rb[n], a);
        r8[mm] = a;
      }
    }
    h_epi_done(s, c.t);
    {
      float acc[64];
      // G0: x_ji = act(lin_ji(e1))                                                spherenet.py:154
      h_drain<4, true>(s, c, col0, 2, acc);
      h_epi_done(s, c.t);   // A (= e1) is reused unchanged by lin_kj: its MMAs may run under this activation
      if (valid) {
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&s.bias[0][col0 + i]);
          float4 o;
          o.x = hswish<FAST>(fmaf(acc[i], H_INV, b.x));
          o.y = hswish<FAST>(fmaf(acc[i + 1], H_INV, b.y));
          o.z = hswish<FAST>(fmaf(acc[i + 2], H_INV, b.z));
          o.w = hswish<FAST>(fmaf(acc[i + 3], H_INV, b.w));
          *reinterpret_cast<float4*>(x_ji + ge * 128 + col0 + i) = o;
        }
      }
      // G1: x_kj = act(lin_kj(e1)) * lin_rbf2(r8)                                 spherenet.py:155-159
      h_drain<4, true>(s, c, col0, 2, acc);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = col0 + 16 * p + i;
          const float4 w0 = *reinterpret_cast<const float4*>(s.wr + col * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(s.wr + col * 8 + 4);
          const float gate = fmaf(w1.w, r8[7], fmaf(w1.z, r8[6], fmaf(w1.y, r8[5], fmaf(w1.x, r8[4],
                             fmaf(w0.w, r8[3], fmaf(w0.z, r8[2], fmaf(w0.y, r8[1], w0.x * r8[0])))))));
          v[i] = hswish8<FAST>(fmaf(acc[16 * p + i], H_SA * H_INV, s.bias[1][col])) * gate;
        }
        h_store_a16(s, c, col0 + 16 * p, v);
      }
      h_epi_done(s, c.t);
    }
    // G2: x_down = act(lin_down(x_kj)), N = 64                                    spherenet.py:161
    {
      const int col = c.half * 32;
      float a32[32];
      h_drain<2, true>(s, c, col, 2, a32);
#pragma unroll
      for (int i = 0; i < 32; ++i) a32[i] = hswish<FAST>(a32[i] * H_INV);
      h_store_tile_coalesced<64>(s, c, col, a32, x_down + (size_t)e0 * 64, rows);   // all MMAs of the tile are done
    }
  }
  h_finish(s, epi ? &c : nullptr);
}

// ---------------------------------------------------------------------------------- init_e
// e1 = act(lin(cat[x_i, x_j, act(lin_rbf_0(rbf))])), e2 = lin_rbf_1(rbf) * e1        spherenet.py:79-91
// K = 384 as three K = 128 panels whose A operand is rebuilt between panels; the chunk sums keep accumulating in
// the epilogue registers across the panels.
struct HInitParams {
  HGemm g[3];
  const float *emb, *w_rbf0, *b_rbf0, *b_lin, *w_rbf1;
  const float *tab_i, *tab_j;   // TABLE: [emb rows, 128] = emb W[:, 0:128]^T and emb W[:, 128:256]^T
};

// TABLE: lin(cat[x_i, x_j, rbf0]) = W_i emb[z_i] + W_j emb[z_j] + W_r rbf0 + b, and the first two terms depend on the
// ATOMIC NUMBER only: they are two [emb rows, 128] tables computed once per parameter version in exact fp32
// (ops.init_e_tables), gathered into the TMEM stash while the single remaining K = 128 panel (the rbf part) runs, and
// added in the final epilogue -- one job per tile instead of three (a job of the two-tile chain is ~10 us, §4.5).
template <bool FAST, bool TABLE>
__global__ void __launch_bounds__(H_THREADS, 1)
sphere_init_e_h16_kernel(const int64_t* __restrict__ z, const int32_t* __restrict__ src,
                         const int32_t* __restrict__ dst, const float* __restrict__ rbf0, int n_edges, HInitParams P,
                         float* __restrict__ e1, float* __restrict__ v_in) {
  extern __shared__ __align__(1024) unsigned char h_raw[];
  HSmem& s = *reinterpret_cast<HSmem*>(h_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (n_edges + H_M - 1) / H_M, tile0 = blockIdx.x * 2, ntile = min(2, n_tiles - tile0);
  float* w0 = &s.bias[2][0];   // lin_rbf_0.weight [128][6] parked in the unused bias rows
  h_setup(s);
  for (int i = tid; i < 128; i += H_THREADS) { s.bias[0][i] = __ldg(P.b_lin + i); s.bias[1][i] = __ldg(P.b_rbf0 + i); }
  for (int i = tid; i < 128 * 6; i += H_THREADS) w0[i] = __ldg(P.w_rbf0 + i);
  for (int i = tid; i < 128 * 8; i += H_THREADS) s.wr[i] = (i % 8 < 6) ? __ldg(P.w_rbf1 + (i / 8) * 6 + i % 8) : 0.f;
  for (int i = tid; i < 2 * H_M; i += H_THREADS) {
    const int t = i / H_M, r = i % H_M, e = (tile0 + t) * H_M + r;
    const int d = (e < n_edges) ? dst[e] : -1, sj = (e < n_edges) ? src[e] : -1;
    s.dst[t][r] = d;
    s.aux[t][0][r] = d >= 0 ? (int)z[d] : 0;
    s.aux[t][1][r] = sj >= 0 ? (int)z[sj] : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  HCtx c;
  bool epi = false;
  if (warp < H_CTRL_WARPS) {
    h_regs_ctrl();
    if (tid == 0) h_producer_n(s, P.g, TABLE ? 1 : 3, ntile);
    else if (tid == 32) h_mma_n<false, false>(s, P.g, TABLE ? 1 : 3, ntile, s.tmem_base);
  } else if (h_regs_epi(), (c = h_ctx(s, ntile)).t < ntile) {
    epi = true;
    const int e0 = (tile0 + c.t) * H_M, rows = min(H_M, n_edges - e0);
    const bool valid = c.row < rows;
    const size_t ge = (size_t)(e0 + c.row);
    const int col0 = c.half * 64;
    const uint32_t stash = c.tl + 256u + 128u * c.t + col0;      // TABLE: the row's table sums wait here
    float rb[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) rb[n] = valid ? __ldg(rbf0 + ge * 6 + n) : 0.f;
    auto fill_embedding = [&](const int* zrow) {   // A = emb[z[node of row]]
#pragma unroll
      for (int k = 0; k < H_M * 16 / H_TILE_THREADS; ++k) {
        const int f = c.et + k * H_TILE_THREADS, row = f >> 4, ku = f & 15;
        float x[8];
        if (row < rows) {
          const float* er = P.emb + (size_t)zrow[row] * 128 + ku * 8;
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(er));
          const float4 p1 = __ldg(reinterpret_cast<const float4*>(er + 4));
          x[0] = p0.x * H_SA; x[1] = p0.y * H_SA; x[2] = p0.z * H_SA; x[3] = p0.w * H_SA;
          x[4] = p1.x * H_SA; x[5] = p1.y * H_SA; x[6] = p1.z * H_SA; x[7] = p1.w * H_SA;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = 0.f;
        }
        h_store_ku(s, c, row, ku, x);
      }
    };
    float acc[64];
    if (!TABLE) {
      fill_embedding(s.aux[c.t][0]);                  // panel 0: x_i
      h_epi_done(s, c.t);
      h_drain<4, true>(s, c, col0, 2, acc);
      fill_embedding(s.aux[c.t][1]);                  // panel 1: x_j
      h_epi_done(s, c.t);
      h_drain<4, false>(s, c, col0, 2, acc);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {                   // panel 2: act(lin_rbf_0(rbf))        spherenet.py:87
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = col0 + 16 * p + i;
        float a = 0.f;
#pragma unroll
        for (int n = 0; n < 6; ++n) a = fmaf(w0[col * 6 + n], rb[n], a);
        v[i] = valid ? H_SA * hswish<FAST>(a + s.bias[1][col]) : 0.f;
      }
      h_store_a16(s, c, col0 + 16 * p, v);
    }
    h_epi_done(s, c.t);
    if (TABLE) {
      // W_i emb[z_i] + W_j emb[z_j] of this row (L2-resident tables) -> stash, under the panel's MMAs
      const float* ti = P.tab_i + (size_t)s.aux[c.t][0][c.row] * 128 + col0;
      const float* tj = P.tab_j + (size_t)s.aux[c.t][1][c.row] * 128 + col0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(ti + 16 * p + i));
          const float4 b = __ldg(reinterpret_cast<const float4*>(tj + 16 * p + i));
          r[i] = __float_as_uint(a.x + b.x); r[i + 1] = __float_as_uint(a.y + b.y);
          r[i + 2] = __float_as_uint(a.z + b.z); r[i + 3] = __float_as_uint(a.w + b.w);
        }
        tmem_st16(stash + 16 * p, r);
      }
      tmem_st_wait();
      h_drain<4, true>(s, c, col0, 2, acc);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        uint32_t r[16];
        tmem_ld16(stash + 16 * p, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[16 * p + i] = fmaf(acc[16 * p + i], H_INV, __uint_as_float(r[i])) * (H_SA * H_SW);
      }
    } else {
      h_drain<4, false>(s, c, col0, 2, acc);
    }
    // e1 = act(. + b), e2 = lin_rbf_1(rbf) * e1 (tile staged over the planes), edge -> node sums, coalesced e1 store
    float* e2t = reinterpret_cast<float*>(s.a[c.t][0]);
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const int col = col0 + i;
      const float o = hswish<FAST>(fmaf(acc[i], H_INV, s.bias[0][col]));
      c.bad |= !h_finite(o);
      const float4 w0 = *reinterpret_cast<const float4*>(s.wr + col * 8);
      const float2 w1 = *reinterpret_cast<const float2*>(s.wr + col * 8 + 4);
      const float gsum = fmaf(w1.y, rb[5], fmaf(w1.x, rb[4], fmaf(w0.w, rb[3], fmaf(w0.z, rb[2],
                         fmaf(w0.y, rb[1], w0.x * rb[0])))));
      e2t[c.row * H_LDS + col] = gsum * o;
      acc[i] = o;
    }
    h_tile_bar(c.t);
    h_segment_sums(s, c, rows, v_in);
    h_tile_bar(c.t);
    h_store_tile_coalesced<128>(s, c, col0, acc, e1 + (size_t)e0 * 128, rows);
  }
  h_finish(s, epi ? &c : nullptr);
}

// ---------------------------------------------------------------------------------- update_v (node MLP)
// v = lin_up(v_in) ; v = act(lins[l](v)) ... ; out = lin(v)                              spherenet.py:212-215
// H = 128 -> O = 256 -> O ... -> out_channels on the same engine in SHARED-A mode: one 128-node tile per CTA whose
// operand (K up to 256) spans both plane regions; the two jobs of a layer are the two 128-column halves of its
// output, drained by the two epilogue groups.  Group X activates its half while the tensor core still runs group
// Y's half; both write the next operand only after every MMA of the layer has completed (l_done).  The last linear
// (out_channels <= 4) is a dot product over the activated row, reduced in a fixed order through shared memory.
// All blocks (init_v + update_vs) run in one launch: blockIdx.y selects the block's weights.
struct HVBlock {
  const unsigned char* p[9];     // packed lin_up [256 x 128], lins[l] [256 x 256] (rows 0..127 then 128..255)
  const float* b[9];             // biases [256]
  const float* w_out;            // [out_channels, 256]
};
struct HVParams { HVBlock blk[8]; int n_lins, out_channels; };

template <bool FAST>
__global__ void __launch_bounds__(H_THREADS, 1)
sphere_update_v_h16_kernel(const float* __restrict__ v_in_all, int n_nodes, HVParams P, float* __restrict__ v_out_all) {
  extern __shared__ __align__(1024) unsigned char h_raw[];
  HSmem& s = *reinterpret_cast<HSmem*>(h_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const HVBlock& B = P.blk[blockIdx.y];
  const int r0 = blockIdx.x * H_M, rows = min(H_M, n_nodes - r0);
  const int ng = P.n_lins + 1, C = P.out_channels;
  __shared__ HGemm g[9];
  h_setup(s, 2 * H_TILE_WARPS);
  if (tid < ng) g[tid] = {B.p[tid], nullptr, tid == 0 ? 128 : 256, 128, (tid == 0 ? 128 : 256) * 128 * 4};
  // biases of the first four layers live in s.bias ([8][128] floats = 4 x 256), later ones are read from global
  for (int i = tid; i < min(ng, 4) * 256; i += H_THREADS) (&s.bias[0][0])[i] = H_SA * __ldg(B.b[i / 256] + i % 256);
  for (int i = tid; i < C * 256; i += H_THREADS) s.wr[i] = __ldg(B.w_out + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < H_CTRL_WARPS) {
    h_regs_ctrl();
    if (tid == 0) h_producer_n(s, g, ng, 2);
    else if (tid == 32) h_mma_n<false, true>(s, g, ng, 2, s.tmem_base);
  } else {
    h_regs_epi();
    HCtx c = h_ctx(s, 2);
    c.ahi = s.a[0][0];
    c.alo = s.a[1][0];
    const int et2 = tid - H_CTRL_THREADS;           // 0 .. 511 over both groups
    const int col0 = c.half * 64;                   // within this group's 128-column half
    const int gcol0 = 128 * c.t + col0;             // output column of acc[0]
    const float* vin = v_in_all + ((size_t)blockIdx.y * n_nodes + r0) * 128;
    // A0 = v_in tile (K = 128), all 512 threads
#pragma unroll
    for (int k = 0; k < H_M * 16 / (2 * H_TILE_THREADS); ++k) {
      const int f = et2 + k * 2 * H_TILE_THREADS, row = f >> 4, ku = f & 15;
      float x[8];
      if (row < rows) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(vin + (size_t)row * 128 + ku * 8));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(vin + (size_t)row * 128 + ku * 8 + 4));
        x[0] = p0.x * H_SA; x[1] = p0.y * H_SA; x[2] = p0.z * H_SA; x[3] = p0.w * H_SA;
        x[4] = p1.x * H_SA; x[5] = p1.y * H_SA; x[6] = p1.z * H_SA; x[7] = p1.w * H_SA;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
      h_store_ku(s, c, row, ku, x);
    }
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&s.a_ready[0]);
    float acc[64];
#pragma unroll 1
    for (int q = 0; q < ng; ++q) {
      h_drain<4, true>(s, c, col0, q == 0 ? 2 : 4, acc);
      // v8 = H_SA * (acc / (H_SA H_SW) + b) ; layers >= 1 apply the activation
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float4 b;
        if (q < 4) b = *reinterpret_cast<const float4*>(&(&s.bias[0][0])[q * 256 + gcol0 + i]);
        else {
          b = __ldg(reinterpret_cast<const float4*>(B.b[q] + gcol0 + i));
          b.x *= H_SA; b.y *= H_SA; b.z *= H_SA; b.w *= H_SA;
        }
        acc[i] = fmaf(acc[i], H_SA * H_INV, b.x); acc[i + 1] = fmaf(acc[i + 1], H_SA * H_INV, b.y);
        acc[i + 2] = fmaf(acc[i + 2], H_SA * H_INV, b.z); acc[i + 3] = fmaf(acc[i + 3], H_SA * H_INV, b.w);
        if (q > 0) {
          acc[i] = hswish8<FAST>(acc[i]); acc[i + 1] = hswish8<FAST>(acc[i + 1]);
          acc[i + 2] = hswish8<FAST>(acc[i + 2]); acc[i + 3] = hswish8<FAST>(acc[i + 3]);
        }
      }
      mbar_wait(&s.l_done, q & 1);          // every MMA that reads the current operand has completed
      tc_fence_after();
      if (q + 1 < ng) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float v16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v16[i] = acc[16 * p + i];
          h_store_a16(s, c, gcol0 + 16 * p, v16);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&s.a_ready[0]);
      }
    }
    // out[row][o] = sum_c v[row][c] * w_out[o][c]: partial sums of this thread's 64 columns, fixed-order reduction
    float* red = reinterpret_cast<float*>(s.a[0][0]);        // [128][4 slots][C], the operand planes are free now
    const int slot = 2 * c.t + c.half;
    for (int o = 0; o < C; ++o) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) part = fmaf(acc[i], s.wr[o * 256 + gcol0 + i], part);
      red[(c.row * 4 + slot) * C + o] = part * (1.0f / H_SA);
      c.bad |= !h_finite(part);
    }
    asm volatile("bar.sync 3, %0;" ::"n"(2 * H_TILE_THREADS) : "memory");
    if (slot == 0 && c.row < rows) {
      float* out = v_out_all + ((size_t)blockIdx.y * n_nodes + r0 + c.row) * C;
      for (int o = 0; o < C; ++o) {
        const float* rr = red + (c.row * 4) * C + o;
        out[o] = ((rr[0] + rr[C]) + rr[2 * C]) + rr[3 * C];
      }
    }
    h_finish(s, &c);
    return;
  }
  h_finish(s, nullptr);
}

// ---------------------------------------------------------------------------------- generic linear (training path)
// y[rows, ldy-strided N columns] = x[rows, K] W^T + bias (+ act_out = swish(y)) on the two-tile engine: K as NPANEL
// panels of K4*8 columns whose operand tile is rebuilt between panels (the chunk sums keep accumulating in the
// epilogue registers), N = 128 or 64 output columns per launch (a wider layer is launched once per 128-column slice).
struct HLinParams { HGemm g[3]; };

template <int NPANEL, int KU, int N>
__global__ void __launch_bounds__(H_THREADS, 1)
linear_h16_kernel(const float* __restrict__ x, int n_rows, int ldx, HLinParams P, const float* __restrict__ bias,
                  float* __restrict__ y, float* __restrict__ act_out, const float* __restrict__ residual, int ldy,
                  int tiles_per_cta, int slice_bytes) {
  extern __shared__ __align__(1024) unsigned char h_raw[];
  HSmem& s = *reinterpret_cast<HSmem*>(h_raw);
  constexpr int NC = N / 2;                                // columns per epilogue thread
  const int tid = threadIdx.x, warp = tid >> 5;
  // blockIdx.y = 128-column slice of a wider layer (all slices of one launch are N wide): its packed weights follow the
  // previous slice's, its bias / output columns start at N * slice.  Small row counts run ONE tile per CTA so that the
  // launch spreads over more SMs (a 4.7 k-row, 256-wide layer is 74 CTAs instead of 19 in two serial launches).
  {
    const int slice = blockIdx.y;
#pragma unroll
    for (int p = 0; p < NPANEL; ++p) P.g[p].w += (size_t)slice * slice_bytes;
    if (bias) bias += slice * N;
    if (y) y += slice * N;
    if (act_out) act_out += slice * N;
    if (residual) residual += slice * N;
  }
  const int n_tiles = (n_rows + H_M - 1) / H_M, tile0 = blockIdx.x * tiles_per_cta;
  const int ntile = min(tiles_per_cta, n_tiles - tile0);
  h_setup(s);
  for (int i = tid; i < N; i += H_THREADS) s.bias[0][i] = bias ? __ldg(bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  HCtx c;
  bool epi = false;
  if (warp < H_CTRL_WARPS) {
    h_regs_ctrl();
    if (tid == 0) h_producer_n(s, P.g, NPANEL, ntile);
    else if (tid == 32) h_mma_n<false, false>(s, P.g, NPANEL, ntile, s.tmem_base);
  } else if (h_regs_epi(), (c = h_ctx(s, ntile)).t < ntile) {
    epi = true;
    const int r0 = (tile0 + c.t) * H_M, rows = min(H_M, n_rows - r0);
    const int col0 = c.half * NC;
    float acc[NC];
#pragma unroll
    for (int p = 0; p < NPANEL; ++p) {
      h_load_tile<KU>(s, c, x + (size_t)r0 * ldx + p * (KU * 8), ldx, rows);
      h_epi_done(s, c.t);
      if (p == 0) h_drain<NC / 16, true>(s, c, col0, KU / 8, acc);
      else h_drain<NC / 16, false>(s, c, col0, KU / 8, acc);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = fmaf(acc[i], H_INV, s.bias[0][col0 + i]);
    // outputs: y = x W^T + b (nullable) and / or act_out = swish(y); `residual` is added to the LAST of them (the skip
    // connection of a residual layer, or the second GEMM of a sum of two linears), y stays the pre-activation
    const float* res = residual ? residual + (size_t)r0 * ldy : nullptr;
    if (y) h_store_tile_strided<N>(s, c, col0, acc, y + (size_t)r0 * ldy, ldy, rows, act_out ? nullptr : res);
    if (act_out) {
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[i] = hswish<false>(acc[i]);
      if (y) h_tile_bar(c.t);
      h_store_tile_strided<N>(s, c, col0, acc, act_out + (size_t)r0 * ldy, ldy, rows, res);
    }
  }
  h_finish(s, epi ? &c : nullptr);
}

// ---------------------------------------------------------------------------------- triplet gather on the tensor cores
// m[e] = sum_{t in trip(e)} x_down[kj(t)] * lin_sbf2(sbf_p[t]) * lin_t2(t_p[t])            spherenet.py:163-171
// The node-centred SIMT kernel (sphere_triplet_gather_node_kernel) is FP32-issue bound: 35 of its ~58 instructions per
// triplet and channel pair are the two 8 -> 64 expansions.  Here those expansions run as tcgen05 MMAs:
//   * one CTA per source node j (all out-edges (j -> i) sum over the same in-edges (k -> j)); the x_down rows of the
//     in-edges are staged in shared memory (padded rows, no L2 gather);
//   * whole out-edges are packed into tiles of <= 128 triplet rows; per tile the projected bases S [128 x 8] and
//     T [128 x 8] are split into fp16 hi / lo planes (K padded to 16 with zeros) and multiplied with the packed
//     lin_sbf2 / lin_t2 weights: G_s, G_t [128 x 64] in TMEM (3 MMAs each, N = 64, K = 16);
//   * epilogue, thread = triplet row: y = x * g_s * g_t into a shared tile, then lane = channel sums the rows of each
//     out-edge and writes m[e] as one 256-byte row.
// The protocol is sequential per tile (build planes -> MMAs -> epilogue): no roles, no rings; two to three CTAs per
// SM overlap each other's phases.
constexpr int GT_ROWS = 128;
constexpr int GT_XLD = 68;                  // floats per staged x_down row (16-byte aligned, spreads the banks)
constexpr int GT_YLD = 65;
constexpr int GT_MAXIN = 64, GT_MAXOUT = 256, GT_MAXSEG = 128;
constexpr float GT_SS = 8.0f, GT_SW = 64.0f;      // operand pre-scales (same rationale as H_SA / H_SW)

struct GTSmem {
  unsigned char s_hi[2 * GT_ROWS * 16], s_lo[2 * GT_ROWS * 16];     // [k-unit][row][8 halves]; k-unit 1 stays zero
  unsigned char t_hi[2 * GT_ROWS * 16], t_lo[2 * GT_ROWS * 16];
  unsigned char ws_hi[2 * 64 * 16], ws_lo[2 * 64 * 16], wt_hi[2 * 64 * 16], wt_lo[2 * 64 * 16];
  float x[GT_MAXIN * GT_XLD];
  float y[GT_ROWS * GT_YLD];
  int in_src[GT_MAXIN];
  int out_e[GT_MAXOUT], out_pos[GT_MAXOUT];
  int seg_row[GT_MAXSEG + 1], seg_out[GT_MAXSEG];   // tile: first row / out-edge slot of each packed out-edge
  int row_seg[GT_ROWS];
  int n_out, n_seg, next_u;
  uint64_t bar;
  uint32_t tmem_base;
};

template <bool TORSION>
__global__ void __launch_bounds__(GT_ROWS)
sphere_triplet_gather_tc_kernel(const float* __restrict__ x_down, const float* __restrict__ sbf_p,
                                const float* __restrict__ t_p, const int32_t* __restrict__ src,
                                const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ trip_ptr,
                                const int32_t* __restrict__ graph_ptr, const int64_t* __restrict__ batch,
                                int n_nodes, const float* __restrict__ w_sbf2, const float* __restrict__ w_t2,
                                float* __restrict__ m) {
  extern __shared__ __align__(1024) unsigned char gt_raw[];
  GTSmem& s = *reinterpret_cast<GTSmem*>(gt_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) { mbar_init(&s.bar, 1); mbar_fence_init(); }
  if (w == 0) tmem_alloc(&s.tmem_base, 128);
  // zero the operand planes once (k-unit 1 and unused rows must read as zero), pack the two small weight matrices
  for (int i = tid; i < (int)(4 * 2 * GT_ROWS * 16 / 16); i += GT_ROWS) reinterpret_cast<uint4*>(s.s_hi)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 64 * 8; i += GT_ROWS) {
    const int n = i >> 3, k = i & 7;
    {
      const float v = __ldg(w_sbf2 + i) * GT_SW;
      const __half h = __float2half_rn(v);
      reinterpret_cast<__half*>(s.ws_hi)[n * 8 + k] = h;
      reinterpret_cast<__half*>(s.ws_lo)[n * 8 + k] = __float2half_rn(v - __half2float(h));
    }
    if (TORSION) {
      const float v = __ldg(w_t2 + i) * GT_SW;
      const __half h = __float2half_rn(v);
      reinterpret_cast<__half*>(s.wt_hi)[n * 8 + k] = h;
      reinterpret_cast<__half*>(s.wt_lo)[n * 8 + k] = __float2half_rn(v - __half2float(h));
    }
  }
  for (int i = tid; i < 64 * 8; i += GT_ROWS) {          // k-unit 1 of the weights: zeros
    reinterpret_cast<__half*>(s.ws_hi)[64 * 8 + i] = __float2half_rn(0.f);
    reinterpret_cast<__half*>(s.ws_lo)[64 * 8 + i] = __float2half_rn(0.f);
    reinterpret_cast<__half*>(s.wt_hi)[64 * 8 + i] = __float2half_rn(0.f);
    reinterpret_cast<__half*>(s.wt_lo)[64 * 8 + i] = __float2half_rn(0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_base;
  const uint32_t idesc = idesc_f16(GT_ROWS, 64);
  uint32_t phase = 0;
  for (int j = blockIdx.x; j < n_nodes; j += gridDim.x) {
    const int base = row_ptr[j], d = row_ptr[j + 1] - base;
    const int g = (int)batch[j], lo = graph_ptr[g], hi = graph_ptr[g + 1];
    __syncthreads();
    // in-neighbour list and x_down rows of j's in-edges
    for (int k = tid; k < d; k += GT_ROWS) s.in_src[k] = src[base + k];
    for (int i = tid; i < d * 16; i += GT_ROWS) {
      const int r = i >> 4, c4 = i & 15;
      *reinterpret_cast<float4*>(&s.x[r * GT_XLD + 4 * c4]) =
          __ldg(reinterpret_cast<const float4*>(x_down + (size_t)(base + r) * 64) + c4);
    }
    for (int c0 = lo; c0 < hi; c0 += GT_MAXOUT) {
      if (tid == 0) s.n_out = 0;
      __syncthreads();
      for (int i = c0 + tid; i < min(hi, c0 + GT_MAXOUT); i += GT_ROWS) {
        if (i == j) continue;
        const int ib = row_ptr[i], di = row_ptr[i + 1] - ib;
        int a = 0, b = di;
        while (a < b) { const int mid = (a + b) >> 1; if (src[ib + mid] < j) a = mid + 1; else b = mid; }
        if (a < di && src[ib + a] == j) {
          int pa = 0, pb = d;
          while (pa < pb) { const int mid = (pa + pb) >> 1; if (s.in_src[mid] < i) pa = mid + 1; else pb = mid; }
          const int slot = atomicAdd(&s.n_out, 1);
          s.out_e[slot] = ib + a;
          s.out_pos[slot] = (pa < d && s.in_src[pa] == i) ? pa : d;
        }
      }
      __syncthreads();
      const int no = s.n_out;
      int u0 = 0;
      while (u0 < no) {                                  // uniform loop: one tile of packed out-edges per iteration
        // pack whole out-edges [u0, u1) while they fit in 128 rows (an out-edge has at most d <= 64 rows)
        if (tid == 0) {
          int rows = 0, u = u0, ns = 0;
          while (u < no && ns < GT_MAXSEG) {
            const int nt = d - (s.out_pos[u] < d ? 1 : 0);
            if (rows + nt > GT_ROWS) break;
            s.seg_row[ns] = rows; s.seg_out[ns] = u;
            rows += nt; ++ns; ++u;
          }
          s.seg_row[ns] = rows;
          s.n_seg = ns;
          s.next_u = u;
        }
        __syncthreads();
        const int ns = s.n_seg, rows = s.seg_row[ns];
        // row -> segment map (warp-parallel fill)
        for (int sg = w; sg < ns; sg += GT_ROWS / 32)
          for (int r = s.seg_row[sg] + lane; r < s.seg_row[sg + 1]; r += 32) s.row_seg[r] = sg;
        __syncthreads();
        // operand planes of this tile: thread = triplet row
        int srow = 0;      // in-edge slot whose x_down row this triplet uses
        {
          const int p = tid;
          float sv[8], tv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sv[i] = tv[i] = 0.f;
          if (p < rows) {
            const int sg = s.row_seg[p], u = s.seg_out[sg], r = p - s.seg_row[sg];
            const int pos_i = s.out_pos[u];
            srow = r + (r >= pos_i ? 1 : 0);
            const size_t t = (size_t)trip_ptr[s.out_e[u]] + r;
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(sbf_p + t * 8));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(sbf_p + t * 8) + 1);
            sv[0] = a0.x; sv[1] = a0.y; sv[2] = a0.z; sv[3] = a0.w; sv[4] = a1.x; sv[5] = a1.y; sv[6] = a1.z; sv[7] = a1.w;
            if (TORSION) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(t_p + t * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(t_p + t * 8) + 1);
              tv[0] = b0.x; tv[1] = b0.y; tv[2] = b0.z; tv[3] = b0.w; tv[4] = b1.x; tv[5] = b1.y; tv[6] = b1.z; tv[7] = b1.w;
            }
          }
          uint32_t h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = sv[2 * i] * GT_SS, b = sv[2 * i + 1] * GT_SS;
            const __half2 hh = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
            h[i] = *reinterpret_cast<const uint32_t*>(&hh); l[i] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          *reinterpret_cast<uint4*>(s.s_hi + p * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(s.s_lo + p * 16) = make_uint4(l[0], l[1], l[2], l[3]);
          if (TORSION) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = tv[2 * i] * GT_SS, b = tv[2 * i + 1] * GT_SS;
              const __half2 hh = __floats2half2_rn(a, b);
              const float2 hf = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
              h[i] = *reinterpret_cast<const uint32_t*>(&hh); l[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            *reinterpret_cast<uint4*>(s.t_hi + p * 16) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(s.t_lo + p * 16) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
          // G_s = S Ws^T into columns [0, 64), G_t = T Wt^T into [64, 128): lo*hi + hi*lo + hi*hi, K = 16 (upper half zero)
          const uint64_t dsh = smem_desc(smem_u32(s.s_hi), GT_ROWS * 16, 128), dsl = smem_desc(smem_u32(s.s_lo), GT_ROWS * 16, 128);
          const uint64_t dwh = smem_desc(smem_u32(s.ws_hi), 64 * 16, 128), dwl = smem_desc(smem_u32(s.ws_lo), 64 * 16, 128);
          mma_f16(tm, dsl, dwh, idesc, 0);
          mma_f16(tm, dsh, dwl, idesc, 1);
          mma_f16(tm, dsh, dwh, idesc, 1);
          if (TORSION) {
            const uint64_t dth = smem_desc(smem_u32(s.t_hi), GT_ROWS * 16, 128), dtl = smem_desc(smem_u32(s.t_lo), GT_ROWS * 16, 128);
            const uint64_t dvh = smem_desc(smem_u32(s.wt_hi), 64 * 16, 128), dvl = smem_desc(smem_u32(s.wt_lo), 64 * 16, 128);
            mma_f16(tm + 64, dtl, dvh, idesc, 0);
            mma_f16(tm + 64, dth, dvl, idesc, 1);
            mma_f16(tm + 64, dth, dvh, idesc, 1);
          }
          mma_commit(&s.bar);
        }
        mbar_wait(&s.bar, phase);
        phase ^= 1;
        tc_fence_after();
        // epilogue: y[p][c] = x[srow][c] * g_s[p][c] * g_t[p][c]
        {
          const uint32_t tl = tm + ((uint32_t)(32 * w) << 16);
          const float scale = TORSION ? 1.0f / (GT_SS * GT_SW * GT_SS * GT_SW) : 1.0f / (GT_SS * GT_SW);
#pragma unroll
          for (int c0c = 0; c0c < 64; c0c += 16) {
            uint32_t gs[16], gt[16];
            tmem_ld16(tl + c0c, gs);
            if (TORSION) tmem_ld16(tl + 64 + c0c, gt);
            float xv[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 q = *reinterpret_cast<const float4*>(&s.x[srow * GT_XLD + c0c + i]);
              xv[i] = q.x * scale; xv[i + 1] = q.y * scale; xv[i + 2] = q.z * scale; xv[i + 3] = q.w * scale;
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float y = xv[i] * __uint_as_float(gs[i]);
              if (TORSION) y *= __uint_as_float(gt[i]);
              s.y[tid * GT_YLD + c0c + i] = y;
            }
          }
        }
        tc_fence_before();
        __syncthreads();
        // segmented sums: one warp per packed out-edge, lane = channel (c, c + 32)
        for (int sg = w; sg < ns; sg += GT_ROWS / 32) {
          const int r0 = s.seg_row[sg], r1 = s.seg_row[sg + 1];
          float a0 = 0.f, a1 = 0.f;
          for (int r = r0; r < r1; ++r) { a0 += s.y[r * GT_YLD + lane]; a1 += s.y[r * GT_YLD + lane + 32]; }
          const size_t e = (size_t)s.out_e[s.seg_out[sg]];
          m[e * 64 + lane] = a0;
          m[e * 64 + lane + 32] = a1;
        }
        u0 = s.next_u;
        __syncthreads();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) tmem_dealloc(s.tmem_base, 128);
}

static int h_smem_attr(const void* fn);
template <int NPANEL, int KU, int N>
static int launch_linear_h16(const float* x, int64_t rows, int ldx, const unsigned char* packed, const float* bias,
                             float* y, float* act_out, const float* residual, int ldy, int slices, cudaStream_t st) {
  HLinParams P;
  const size_t panel = (size_t)(KU / 4) * 2 * 4 * N * 16;      // KU/4 slabs of [hi|lo][4][N][8 halves]
  for (int p = 0; p < NPANEL; ++p) P.g[p] = {packed + p * panel, nullptr, KU * 8, N};
  auto kfn = linear_h16_kernel<NPANEL, KU, N>;
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ceil_div(rows, H_M);
  const int tpc = (ceil_div(tiles, 2) * slices < n_sm) ? 1 : 2;        // below one wave of tile pairs: one tile per CTA
  dim3 grid(ceil_div(tiles, tpc), slices);
  kfn<<<grid, H_THREADS, sizeof(HSmem), st>>>(x, (int)rows, ldx, P, bias, y, act_out, residual, ldy, tpc,
                                              4 * N * (NPANEL * KU * 8));
  return DIG3D_OK;
}

static int h_smem_attr(const void* fn) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HSmem));
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu): %s", sizeof(HSmem), cudaGetErrorString(e));
    return DIG3D_ECUDA;
  }
  return DIG3D_OK;
}

}  // namespace dig3d

using namespace dig3d;

extern "C" {

int64_t dig3d_h16_packed_bytes(int32_t n, int32_t k) { return (int64_t)4 * n * k; }

static int h16_pack_impl(const float* const* weights, const int32_t* n, const int32_t* k, const int32_t* trans,
                         void* const* outs, int32_t count, void* stream) {
  DIG3D_REQUIRE(weights && n && k && outs && count >= 1 && count <= 16, "h16_pack: bad arguments");
  HPackJobs jobs;
  int max_total = 0;
  for (int i = 0; i < count; ++i) {
    DIG3D_REQUIRE(weights[i] && outs[i] && k[i] % 64 == 0 && n[i] % 8 == 0 && n[i] <= 128,
                  "h16_pack: matrix %d has N=%d K=%d (need K %% 64 == 0, N %% 8 == 0, N <= 128)", i, n[i], k[i]);
    jobs.job[i] = {weights[i], (unsigned char*)outs[i], n[i], k[i], trans ? trans[i] : 0};
    max_total = max_total > n[i] * k[i] ? max_total : n[i] * k[i];
  }
  dim3 grid(ceil_div(max_total, 256), count);
  h16_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(jobs);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_h16_pack(const float* const* weights, const int32_t* n, const int32_t* k, void* const* outs, int32_t count,
                   void* stream) {
  return h16_pack_impl(weights, n, k, nullptr, outs, count, stream);
}

int dig3d_h16_pack_t(const float* const* weights, const int32_t* n, const int32_t* k, const int32_t* trans,
                     void* const* outs, int32_t count, void* stream) {
  DIG3D_REQUIRE(trans, "h16_pack_t: null pointer");
  return h16_pack_impl(weights, n, k, trans, outs, count, stream);
}

int dig3d_h16_set_fast_swish(int32_t on) {
  h16_fast_swish = on ? 1 : 0;
  return DIG3D_OK;
}

int dig3d_h16_set_wide_epilogue(int32_t on) {
  h16_wide_epilogue = on ? 1 : 0;
  return DIG3D_OK;
}

int dig3d_h16_overflow(int32_t clear) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_h16_overflow, sizeof(v));
  if (clear && v) {
    unsigned int zero = 0;
    cudaMemcpyToSymbol(g_h16_overflow, &zero, sizeof(zero));
  }
  return (int)v;
}

int dig3d_h16_trace(int32_t on, long long* out128 /* host, 128 entries, nullable */) {
  if (out128) cudaMemcpyFromSymbol(out128, g_h16_trace, sizeof(long long) * 128);
  int v = on ? 1 : 0;
  cudaMemcpyToSymbol(g_h16_trace_on, &v, sizeof(v));
  return DIG3D_OK;
}

int dig3d_h16_timeouts(void) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, tc05::g_mbar_timeout, sizeof(v));
  return (int)v;
}

int dig3d_sphere_init_e_h16(const int64_t* z, const int32_t* src, const int32_t* dst, const float* rbf0,
                            int64_t n_edges, const dig3d_init_e_weights* w, const void* packed_lin, float* e1,
                            float* v_in, void* stream) {
  DIG3D_REQUIRE(z && src && dst && rbf0 && w && packed_lin && e1 && v_in, "sphere_init_e_h16: null pointer");
  DIG3D_REQUIRE(w->emb && w->w_rbf0 && w->b_rbf0 && w->b_lin && w->w_rbf1, "sphere_init_e_h16: null weight");
  if (n_edges == 0) return DIG3D_OK;
  HInitParams P;
  const size_t panel = (size_t)4 * 2 * 4 * 128 * 16;   // four K=32 slabs
  for (int p = 0; p < 3; ++p) P.g[p] = {(const unsigned char*)packed_lin + p * panel, nullptr, 128, 128};
  P.emb = w->emb; P.w_rbf0 = w->w_rbf0; P.b_rbf0 = w->b_rbf0; P.b_lin = w->b_lin; P.w_rbf1 = w->w_rbf1;
  P.tab_i = P.tab_j = nullptr;
  auto kfn = h16_fast_swish ? sphere_init_e_h16_kernel<true, false> : sphere_init_e_h16_kernel<false, false>;
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  const int pairs = ceil_div(ceil_div(n_edges, H_M), 2);
  kfn<<<pairs, H_THREADS, sizeof(HSmem), (cudaStream_t)stream>>>(z, src, dst, rbf0, (int)n_edges, P, e1, v_in);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_init_e_h16_tab(const int64_t* z, const int32_t* src, const int32_t* dst, const float* rbf0,
                                int64_t n_edges, const dig3d_init_e_weights* w, const void* packed_rbf_panel,
                                const float* tab_i, const float* tab_j, float* e1, float* v_in, void* stream) {
  DIG3D_REQUIRE(z && src && dst && rbf0 && w && packed_rbf_panel && tab_i && tab_j && e1 && v_in,
                "sphere_init_e_h16_tab: null pointer");
  DIG3D_REQUIRE(w->w_rbf0 && w->b_rbf0 && w->b_lin && w->w_rbf1, "sphere_init_e_h16_tab: null weight");
  DIG3D_REQUIRE((((uintptr_t)tab_i | (uintptr_t)tab_j) & 15) == 0, "sphere_init_e_h16_tab: tables must be 16-byte aligned");
  if (n_edges == 0) return DIG3D_OK;
  HInitParams P;
  P.g[0] = {(const unsigned char*)packed_rbf_panel, nullptr, 128, 128};
  P.g[1] = P.g[2] = P.g[0];
  P.emb = w->emb; P.w_rbf0 = w->w_rbf0; P.b_rbf0 = w->b_rbf0; P.b_lin = w->b_lin; P.w_rbf1 = w->w_rbf1;
  P.tab_i = tab_i; P.tab_j = tab_j;
  auto kfn = h16_fast_swish ? sphere_init_e_h16_kernel<true, true> : sphere_init_e_h16_kernel<false, true>;
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  const int pairs = ceil_div(ceil_div(n_edges, H_M), 2);
  kfn<<<pairs, H_THREADS, sizeof(HSmem), (cudaStream_t)stream>>>(z, src, dst, rbf0, (int)n_edges, P, e1, v_in);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_e_a_h16(const float* e1, const float* rbf0, int64_t n_edges, const dig3d_tc_update_e* w,
                                float* x_ji, float* x_down, void* stream) {
  DIG3D_REQUIRE(e1 && rbf0 && w && x_ji && x_down, "sphere_update_e_a_h16: null pointer");
  if (n_edges == 0) return DIG3D_OK;
  HAParams P;
  P.g[0] = {(const unsigned char*)w->p_ji, w->b_ji, 128, 128};
  P.g[1] = {(const unsigned char*)w->p_kj, w->b_kj, 128, 128};
  P.g[2] = {(const unsigned char*)w->p_down, nullptr, 128, 64};
  P.w_rbf1 = w->w_rbf1; P.w_rbf2 = w->w_rbf2;
  auto kfn = h16_fast_swish ? sphere_update_e_a_h16_kernel<true> : sphere_update_e_a_h16_kernel<false>;
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  const int pairs = ceil_div(ceil_div(n_edges, H_M), 2);
  kfn<<<pairs, H_THREADS, sizeof(HSmem), (cudaStream_t)stream>>>(e1, rbf0, (int)n_edges, P, x_ji, x_down);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_e_b_h16(const float* m, const float* e1_in, const float* x_ji, const float* rbf0,
                                const int32_t* dst, int64_t n_edges, const dig3d_tc_update_e* w, float* e1_out,
                                float* v_in, void* stream) {
  DIG3D_REQUIRE(m && e1_in && x_ji && rbf0 && dst && w && e1_out && v_in, "sphere_update_e_b_h16: null pointer");
  if (n_edges == 0) return DIG3D_OK;
  HBParams P;
  P.g[0] = {(const unsigned char*)w->p_up, nullptr, 64, 128};
  for (int i = 0; i < 2; ++i) P.g[1 + i] = {(const unsigned char*)w->p_res[i], w->b_res[i], 128, 128};
  P.g[3] = {(const unsigned char*)w->p_lin, w->b_lin, 128, 128};
  for (int i = 2; i < 6; ++i) P.g[2 + i] = {(const unsigned char*)w->p_res[i], w->b_res[i], 128, 128};
  P.w_rbf = w->w_rbf;
  P.n_w_rbf1 = P.n_w_rbf2 = nullptr; P.n_x_ji = P.n_x_down = nullptr;
  auto kfn = h16_wide_epilogue
                 ? (h16_fast_swish ? sphere_update_e_b_x16_kernel<true, false> : sphere_update_e_b_x16_kernel<false, false>)
                 : (h16_fast_swish ? sphere_update_e_b_h16_kernel<true, false> : sphere_update_e_b_h16_kernel<false, false>);
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  const int pairs = ceil_div(ceil_div(n_edges, H_M), 2);
  kfn<<<pairs, H_THREADS, sizeof(HSmem), (cudaStream_t)stream>>>(m, x_ji, e1_in, rbf0, dst, (int)n_edges, P, e1_out,
                                                                 v_in);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_e_ba_h16(const float* m, const float* e1_in, const float* x_ji, const float* rbf0,
                                 const int32_t* dst, int64_t n_edges, const dig3d_tc_update_e* w,
                                 const dig3d_tc_update_e* w_next, float* e1_out, float* v_in, float* x_ji_next,
                                 float* x_down_next, void* stream) {
  DIG3D_REQUIRE(m && e1_in && x_ji && rbf0 && dst && w && w_next && e1_out && v_in && x_ji_next && x_down_next,
                "sphere_update_e_ba_h16: null pointer");
  DIG3D_REQUIRE(x_ji_next != x_ji, "sphere_update_e_ba_h16: x_ji_next must not alias x_ji (other tiles still read it)");
  if (n_edges == 0) return DIG3D_OK;
  HBParams P;
  P.g[0] = {(const unsigned char*)w->p_up, nullptr, 64, 128};
  for (int i = 0; i < 2; ++i) P.g[1 + i] = {(const unsigned char*)w->p_res[i], w->b_res[i], 128, 128};
  P.g[3] = {(const unsigned char*)w->p_lin, w->b_lin, 128, 128};
  for (int i = 2; i < 6; ++i) P.g[2 + i] = {(const unsigned char*)w->p_res[i], w->b_res[i], 128, 128};
  P.g[8] = {(const unsigned char*)w_next->p_ji, w_next->b_ji, 128, 128};
  P.g[9] = {(const unsigned char*)w_next->p_kj, w_next->b_kj, 128, 128};
  P.g[10] = {(const unsigned char*)w_next->p_down, nullptr, 128, 64};
  P.w_rbf = w->w_rbf;
  P.n_w_rbf1 = w_next->w_rbf1; P.n_w_rbf2 = w_next->w_rbf2;
  P.n_x_ji = x_ji_next; P.n_x_down = x_down_next;
  auto kfn = h16_wide_epilogue
                 ? (h16_fast_swish ? sphere_update_e_b_x16_kernel<true, true> : sphere_update_e_b_x16_kernel<false, true>)
                 : (h16_fast_swish ? sphere_update_e_b_h16_kernel<true, true> : sphere_update_e_b_h16_kernel<false, true>);
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  const int pairs = ceil_div(ceil_div(n_edges, H_M), 2);
  kfn<<<pairs, H_THREADS, sizeof(HSmem), (cudaStream_t)stream>>>(m, x_ji, e1_in, rbf0, dst, (int)n_edges, P, e1_out,
                                                                 v_in);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_triplet_gather_tc(const float* x_down, const float* sbf_p, const float* t_p, int32_t ld_p,
                                   const int32_t* src, const int32_t* row_ptr, const int32_t* trip_ptr,
                                   const int32_t* graph_ptr, const int64_t* batch, int64_t n_nodes, int32_t cap,
                                   const float* w_sbf2, const float* w_t2, float* m, void* stream) {
  DIG3D_REQUIRE(x_down && sbf_p && src && row_ptr && trip_ptr && graph_ptr && batch && w_sbf2 && m,
                "sphere_triplet_gather_tc: null pointer");
  DIG3D_REQUIRE((t_p != nullptr) == (w_t2 != nullptr), "sphere_triplet_gather_tc: t_p and w_t2 must agree");
  DIG3D_REQUIRE(ld_p == 8, "sphere_triplet_gather_tc: expects the layer-major [T, 8] slices (ld_p == 8), got %d", ld_p);
  DIG3D_REQUIRE(cap >= 1 && cap <= GT_MAXIN, "sphere_triplet_gather_tc: cap=%d outside [1,%d]", cap, GT_MAXIN);
  DIG3D_REQUIRE((((uintptr_t)x_down | (uintptr_t)sbf_p | (uintptr_t)t_p) & 15) == 0, "sphere_triplet_gather_tc: 16-byte alignment");
  if (n_nodes == 0) return DIG3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_nodes < 2 * n_sm ? n_nodes : 2 * n_sm);   // 79 KB of shared memory per CTA: two per SM
  cudaError_t e;
  if (t_p) {
    e = cudaFuncSetAttribute(sphere_triplet_gather_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GTSmem));
    if (e == cudaSuccess)
      sphere_triplet_gather_tc_kernel<true><<<grid, GT_ROWS, sizeof(GTSmem), st>>>(
          x_down, sbf_p, t_p, src, row_ptr, trip_ptr, graph_ptr, batch, (int)n_nodes, w_sbf2, w_t2, m);
  } else {
    e = cudaFuncSetAttribute(sphere_triplet_gather_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GTSmem));
    if (e == cudaSuccess)
      sphere_triplet_gather_tc_kernel<false><<<grid, GT_ROWS, sizeof(GTSmem), st>>>(
          x_down, sbf_p, t_p, src, row_ptr, trip_ptr, graph_ptr, batch, (int)n_nodes, w_sbf2, w_t2, m);
  }
  if (e != cudaSuccess) { set_error("sphere_triplet_gather_tc: %s", cudaGetErrorString(e)); return DIG3D_ECUDA; }
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_linear_h16_supported(int32_t k, int32_t nout) {
  return ((k == 64 || k == 128 || k == 256 || k == 384) && nout >= 64 && nout % 64 == 0 && nout <= 512) ? 1 : 0;
}

/* y = x W^T + bias (+ act_out = swish(y)); packed = dig3d_h16_pack(_t) of W as consecutive [min(128, N - c), K] row
 * slices (one per 128 output columns; a trailing 64-column slice is allowed). */
int dig3d_linear_h16(const float* x, int64_t rows, int32_t k, int32_t nout, const void* packed, const float* bias,
                     float* y, float* act_out, const float* residual, void* stream) {
  DIG3D_REQUIRE(x && packed && (y || act_out), "linear_h16: null pointer");
  DIG3D_REQUIRE(dig3d_linear_h16_supported(k, nout), "linear_h16: shape %d -> %d is not compiled", k, nout);
  DIG3D_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)act_out | (uintptr_t)residual) & 15) == 0,
                "linear_h16: 16-byte alignment");
  if (rows == 0) return DIG3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned char* pw = (const unsigned char*)packed;
  // all full 128-column slices in ONE launch (gridDim.y); a trailing 64-column slice gets its own
  for (int c0 = 0; c0 < nout;) {
    const int n = nout - c0 >= 128 ? 128 : 64;
    const int slices = n == 128 ? (nout - c0) / 128 : 1;
    const float* b = bias ? bias + c0 : nullptr;
    float* yo = y ? y + c0 : nullptr;
    float* ao = act_out ? act_out + c0 : nullptr;
    const float* ro = residual ? residual + c0 : nullptr;
    int rc;
#define DIG3D_LIN(NP, KU)                                                                                             \
    (n == 128 ? launch_linear_h16<NP, KU, 128>(x, rows, k, pw, b, yo, ao, ro, nout, slices, st)                        \
              : launch_linear_h16<NP, KU, 64>(x, rows, k, pw, b, yo, ao, ro, nout, slices, st))
    if (k == 64) rc = DIG3D_LIN(1, 8);
    else if (k == 128) rc = DIG3D_LIN(1, 16);
    else if (k == 256) rc = DIG3D_LIN(2, 16);
    else rc = DIG3D_LIN(3, 16);
#undef DIG3D_LIN
    if (rc) return rc;
    pw += (size_t)4 * n * k * slices;
    c0 += n * slices;
  }
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_sphere_update_v_h16_supported(int32_t hidden, int32_t out_emb, int32_t out_channels, int32_t n_lins) {
  return (hidden == 128 && out_emb == 256 && out_channels >= 1 && out_channels <= 4 && n_lins >= 0 && n_lins <= 8) ? 1 : 0;
}

int dig3d_sphere_update_v_h16(const float* v_in_all, int64_t n_nodes, int32_t n_blocks, int32_t out_channels,
                              int32_t n_lins, const void* const* packed /* [n_blocks][n_lins + 1] */,
                              const dig3d_update_v_weights* w, float* v_out_all, void* stream) {
  DIG3D_REQUIRE(v_in_all && packed && w && v_out_all, "sphere_update_v_h16: null pointer");
  DIG3D_REQUIRE(n_blocks >= 1 && n_blocks <= 8, "sphere_update_v_h16: n_blocks=%d outside [1,8]", n_blocks);
  DIG3D_REQUIRE(dig3d_sphere_update_v_h16_supported(128, 256, out_channels, n_lins),
                "sphere_update_v_h16: out_channels=%d / n_lins=%d not compiled", out_channels, n_lins);
  if (n_nodes == 0) return DIG3D_OK;
  HVParams P;
  P.n_lins = n_lins; P.out_channels = out_channels;
  for (int b = 0; b < n_blocks; ++b) {
    DIG3D_REQUIRE(w[b].n_lins == n_lins && w[b].w_out && w[b].b_up, "sphere_update_v_h16: block %d weights", b);
    for (int l = 0; l <= n_lins; ++l) {
      P.blk[b].p[l] = (const unsigned char*)packed[b * (n_lins + 1) + l];
      P.blk[b].b[l] = l == 0 ? w[b].b_up : w[b].b_lins[l - 1];
      DIG3D_REQUIRE(P.blk[b].p[l] && P.blk[b].b[l], "sphere_update_v_h16: block %d layer %d null", b, l);
    }
    P.blk[b].w_out = w[b].w_out;
  }
  auto kfn = h16_fast_swish ? sphere_update_v_h16_kernel<true> : sphere_update_v_h16_kernel<false>;
  int rc = h_smem_attr((const void*)kfn);
  if (rc) return rc;
  dim3 grid(ceil_div(n_nodes, H_M), n_blocks);
  kfn<<<grid, H_THREADS, sizeof(HSmem), (cudaStream_t)stream>>>(v_in_all, (int)n_nodes, P, v_out_all);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

}  // extern "C"
