// Weight gradient on the tensor cores (training path):  dW[n][k] += sum_r dY[r][n] X[r][k],  db[n] += sum_r dY[r][n]
// (reference: torch.autograd of nn.Linear inside run.py:118 `loss.backward()`; spherenet.py:79-216 holds the layers).
//
// The reduction runs over the ROWS (edges), so both operands of the GEMM  D[n][k] = A[n][r] B[k][r]^T  are the
// transposes of row-major activations.  tcgen05 wants K-major operand tiles ([k-unit][row][4 tf32]); the producer
// warps build them on the fly: lane = column (coalesced 128-byte reads of four consecutive rows), the four values are
// one 16-byte k-unit entry of that column -> one conflict-free st.shared.v4 per plane.  No transposed copy of the
// activations is ever materialised.
//   * 3xTF32 (A_lo B_hi + A_hi B_lo + A_hi B_hi): fp32-level accuracy with fp32 exponent range (gradients span many
//     decades; the 3xFP16 split of the forward engine would need a per-tensor scale);
//   * 32 rows per stage, 3 stages; chunk c accumulates into TMEM accumulator c % 4 (the tensor-core accumulate
//     truncates, so the chain of additions per accumulator is kept short), the epilogue adds the four in registers;
//   * grid = (K tiles, N tiles, row splits x groups) ~ one CTA per SM; partial tiles are staged through shared memory and
//     added to dW with coalesced red.global (dW / db are zero-initialised by the caller, as for the FFMA kernel).
#include "common.cuh"
#include "tc05.cuh"

namespace dig3d {
using namespace tc05;

constexpr int WG_T = 128;                 // tile edge (n and k)
constexpr int WG_RC = 32;                 // rows per stage
constexpr int WG_STAGES = 3;
constexpr int WG_PROD = 256;              // producer / epilogue threads (8 warps)
constexpr int WG_THREADS = WG_PROD + 32;  // + the MMA warp
constexpr int WG_PLANE = (WG_RC / 4) * WG_T * 4;     // floats per operand plane and stage: [8 k-units][128][4]
constexpr int WG_TLD = WG_T + 1;

struct WgSmem {
  float plane[WG_STAGES][4][WG_PLANE];    // a_hi, a_lo, b_hi, b_lo                                  (3 x 64 KB)
  uint64_t full[WG_STAGES], empty[WG_STAGES], done;
  uint32_t tmem_base;
};
static_assert(WG_T * WG_TLD * 4 <= 2 * 4 * WG_PLANE * 4, "the epilogue tile reuses the first two stages");

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t rows, int nout, int k,
                float* __restrict__ dw, float* __restrict__ db, int splits) {
  extern __shared__ __align__(1024) unsigned char wg_raw[];
  WgSmem& s = *reinterpret_cast<WgSmem*>(wg_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.z % splits;
  {
    const size_t gi = blockIdx.z / splits;
    dy += gi * (size_t)rows * nout;
    x += gi * (size_t)rows * k;
    dw += gi * (size_t)nout * k;
    if (db) db += gi * nout;
  }
  const int kb = blockIdx.x * WG_T, nb = blockIdx.y * WG_T;
  int64_t per = (rows + splits - 1) / splits;
  per = (per + WG_RC - 1) / WG_RC * WG_RC;
  const int64_t r_lo = (int64_t)split * per, r_hi = min(rows, r_lo + per);
  if (r_lo >= r_hi) return;                                     // uniform over the CTA, before any allocation
  const int n_chunks = (int)((r_hi - r_lo + WG_RC - 1) / WG_RC);
  if (tid == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&s.full[i], WG_PROD / 32); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.done, 1);
    mbar_fence_init();
  }
  if (warp == WG_PROD / 32) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_base;
  if (warp < WG_PROD / 32) {
    // ---------------- producers: column c of both operands, k-units half, half + 2, half + 4, half + 6
    const int c = tid & (WG_T - 1), half = tid >> 7;
    const bool n_ok = nb + c < nout, k_ok = kb + c < k;
    const bool do_bias = db != nullptr && blockIdx.x == 0 && n_ok;
    const float* dyc = dy + nb + c;
    const float* xc = x + kb + c;
    float bsum = 0.f;
    float av[4][4], bv[4][4];
    // this thread's 16 + 16 values of chunk ch (zero beyond the slice): running pointers, one 64-bit add per load; the
    // bounds are only tested in the (single) ragged last chunk
    const size_t step_a = (size_t)nout, step_b = (size_t)k;
    auto fetch = [&](int ch) {
      const int64_t r0 = r_lo + (int64_t)ch * WG_RC + half * 4;
      const float* pa = dyc + r0 * nout;
      const float* pb = xc + r0 * k;
      if (r_lo + (int64_t)(ch + 1) * WG_RC <= r_hi) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            av[u][q] = n_ok ? __ldg(pa) : 0.f;
            bv[u][q] = k_ok ? __ldg(pb) : 0.f;
            pa += step_a; pb += step_b;
          }
          pa += 4 * step_a; pb += 4 * step_b;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool in = r0 + u * 8 + q < r_hi;
            av[u][q] = (in && n_ok) ? __ldg(pa) : 0.f;
            bv[u][q] = (in && k_ok) ? __ldg(pb) : 0.f;
            pa += step_a; pb += step_b;
          }
          pa += 4 * step_a; pb += 4 * step_b;
        }
      }
    };
    fetch(0);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int st = ch % WG_STAGES;
      if (ch >= WG_STAGES) mbar_wait(&s.empty[st], ((ch / WG_STAGES) - 1) & 1);
      float4 ah[4], al[4], bh[4], bl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        split_tf32(av[u][0], ah[u].x, al[u].x); split_tf32(av[u][1], ah[u].y, al[u].y);
        split_tf32(av[u][2], ah[u].z, al[u].z); split_tf32(av[u][3], ah[u].w, al[u].w);
        split_tf32(bv[u][0], bh[u].x, bl[u].x); split_tf32(bv[u][1], bh[u].y, bl[u].y);
        split_tf32(bv[u][2], bh[u].z, bl[u].z); split_tf32(bv[u][3], bh[u].w, bl[u].w);
        bsum += (av[u][0] + av[u][1]) + (av[u][2] + av[u][3]);
      }
      if (ch + 1 < n_chunks) fetch(ch + 1);          // the next chunk's loads fly while this one is stored and multiplied
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int o = ((half + 2 * u) * WG_T + c) * 4;
        *reinterpret_cast<float4*>(&s.plane[st][0][o]) = ah[u];
        *reinterpret_cast<float4*>(&s.plane[st][1][o]) = al[u];
        *reinterpret_cast<float4*>(&s.plane[st][2][o]) = bh[u];
        *reinterpret_cast<float4*>(&s.plane[st][3][o]) = bl[u];
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.full[st]);
    }
    if (do_bias) atomicAdd(db + nb + c, bsum);
  } else if (lane == 0) {
    // ---------------- MMA issuer
    const uint32_t idesc = idesc_tf32(WG_T, WG_T);
    constexpr uint32_t LBO = WG_T * 16, SBO = 128;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int st = ch % WG_STAGES;
      mbar_wait(&s.full[st], (ch / WG_STAGES) & 1);
      tc_fence_after();
      const uint32_t d = tm + (uint32_t)(ch & 3) * WG_T;
      const uint32_t a_hi = smem_u32(s.plane[st][0]), a_lo = smem_u32(s.plane[st][1]);
      const uint32_t b_hi = smem_u32(s.plane[st][2]), b_lo = smem_u32(s.plane[st][3]);
#pragma unroll
      for (int ks = 0; ks < WG_RC / 8; ++ks) {                   // K = 8 rows (two k-units) per instruction
        const uint32_t off = ks * 2 * LBO;
        const uint32_t first = (ch < 4 && ks == 0) ? 0u : 1u;
        mma_tf32(d, smem_desc(a_lo + off, LBO, SBO), smem_desc(b_hi + off, LBO, SBO), idesc, first);
        mma_tf32(d, smem_desc(a_hi + off, LBO, SBO), smem_desc(b_lo + off, LBO, SBO), idesc, 1u);
        mma_tf32(d, smem_desc(a_hi + off, LBO, SBO), smem_desc(b_hi + off, LBO, SBO), idesc, 1u);
      }
      mma_commit(&s.empty[st]);
    }
    mma_commit(&s.done);
  }
  // ---------------- epilogue: TMEM -> shared tile -> coalesced red.global
  float* tile = &s.plane[0][0][0];
  if (warp < WG_PROD / 32) {
    mbar_wait(&s.done, 0);
    tc_fence_after();
    const int q = warp & 3, hcol = (warp >> 2) * 64;
    const int n = 32 * q + lane;
    const uint32_t tl = tm + ((uint32_t)(32 * q) << 16);
    const int nacc = n_chunks < 4 ? n_chunks : 4;
#pragma unroll 1
    for (int c0 = hcol; c0 < hcol + 64; c0 += 16) {
      float acc[16], t[16];
      tmem_ld16f(tl + c0, acc);
      tmem_ld_wait();
      for (int a = 1; a < nacc; ++a) {
        tmem_ld16f(tl + a * WG_T + c0, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) tile[n * WG_TLD + c0 + i] = acc[i];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp < WG_PROD / 32) {
    for (int n = warp; n < WG_T; n += WG_PROD / 32) {
      if (nb + n >= nout) break;
      float* row = dw + (size_t)(nb + n) * k + kb;
#pragma unroll
      for (int j = 0; j < WG_T / 32; ++j) {
        const int kk = lane + 32 * j;
        if (kb + kk < k) red_add(row + kk, tile[n * WG_TLD + kk]);
      }
    }
  }
  __syncthreads();
  if (warp == WG_PROD / 32) tmem_dealloc(tm, 512);
}

static int h_wgrad_mode = 1;   // 1: tensor cores where the shape pays, 0: FFMA kernel only

}  // namespace dig3d

using namespace dig3d;

extern "C" {

int dig3d_wgrad_set_mode(int32_t mode) {
  DIG3D_REQUIRE(mode == 0 || mode == 1, "wgrad_set_mode: mode must be 0 (FFMA) or 1 (tcgen05 3xTF32)");
  h_wgrad_mode = mode;
  return DIG3D_OK;
}

int dig3d_wgrad_tc_supported(int64_t rows, int32_t nout, int32_t k) {
  return (h_wgrad_mode == 1 && nout >= 64 && k >= 64 && rows >= 1024) ? 1 : 0;
}

int dig3d_wgrad_tc(const float* dy, const float* x, int64_t rows, int32_t nout, int32_t k, float* dw, float* db,
                   int32_t groups, void* stream) {
  DIG3D_REQUIRE(dy && x && dw && nout > 0 && k > 0 && groups >= 1 && groups <= 64, "wgrad_tc: bad arguments");
  if (rows == 0) return DIG3D_OK;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ceil_div(k, WG_T) * ceil_div(nout, WG_T) * groups;
  int64_t splits = n_sm / tiles;
  const int64_t max_splits = (rows + 2 * WG_RC - 1) / (2 * WG_RC);       // at least two stages of rows per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WgSmem));
  if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute(%zu): %s", sizeof(WgSmem), cudaGetErrorString(e)); return DIG3D_ECUDA; }
  dim3 grid(ceil_div(k, WG_T), ceil_div(nout, WG_T), (unsigned)(splits * groups));
  wgrad_tc_kernel<<<grid, WG_THREADS, sizeof(WgSmem), (cudaStream_t)stream>>>(dy, x, rows, nout, k, dw, db, (int)splits);
  DIG3D_LAUNCH_CHECK();
  return DIG3D_OK;
}

int dig3d_wgrad_tc_timeouts(void) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, tc05::g_mbar_timeout, sizeof(v));
  return (int)v;
}

}  // extern "C"
