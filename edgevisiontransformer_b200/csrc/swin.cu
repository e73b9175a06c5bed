// Kernels of the Swin shifted-window block (the model tools.py:265-292 export_onnx_swin /
// utils.py:14-47 get_swin builds; arithmetic = SwinLayer / SwinSelfAttention / SwinPatchMerging,
// SITE/models/swin/modeling_swin.py:141-160, 326-349, 410-459, 591-654).
//
// Layout decision: inside a stage the token rows are kept in WINDOW ORDER of the current block (image, window, token
// in window), so the QKV / out-proj / FC1 / FC2 GEMMs and their TMA reduce-add epilogues work on plain row-major
// matrices and a window's 49 tokens are 49 consecutive rows.  Changing the order (cyclic shift on / off, raster ->
// windows after the patch embedding, the 2x2 neighbourhood gather of patch merging) is a row gather fused into the
// LayerNorm that follows it anyway:
//
//   gather_ln_kernel          out row r = LayerNorm( concat_g x[ image(r) * T_in + idx[(r % T_out) * G + g] ] ), G = 1 | 4;
//                             optionally also the raw gathered row as f32 (the permuted residual stream).  HBM-bound.
//   window_attention_kernel   softmax(q k^T * scale + bias_table) v for 49-token windows, head size 32, one warp per
//                             (window, head): q, k, v staged with cp.async into swizzled shared memory, mma.sync
//                             m16n8k16 bf16 with the probabilities kept in registers (accumulator fragments re-used as A
//                             fragments).  A 49 x 49 x 32 problem fills 38 % of a 128-row tcgen05 tile and its time is
//                             exp / load bound, so the warp-level tensor path is used here; packing two windows per
//                             tcgen05 tile is the follow-up.
//   ln_mean_tokens_kernel     final LayerNorm + average pool over the tokens of an image (SwinModel.layernorm + pooler).
#include <cuda_bf16.h>
#include <math_constants.h>

#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace evt {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------- gather + LN
// One warp per output row of D = G * C floats (C % 4 == 0, D <= 128 * NV).
template <int NV, bool OUT_BF16>
__global__ void __launch_bounds__(256) gather_ln_kernel(const float* __restrict__ x, const int* __restrict__ idx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        void* __restrict__ y, float* __restrict__ copy, long long rows_out,
                                                        int T_in, int T_out, int G, int C, float eps) {
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows_out) return;
  const int D = G * C;
  const long long img = row / T_out;
  const int t = static_cast<int>(row - img * T_out);
  const float* xb = x + img * T_in * static_cast<long long>(C);
  const int* ir = idx + static_cast<long long>(t) * G;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
      const int g = c / C;
      const int src = __ldg(ir + g);
      v[i] = *reinterpret_cast<const float4*>(xb + static_cast<long long>(src) * C + (c - g * C));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  if (copy != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < D) *reinterpret_cast<float4*>(copy + row * D + c) = v[i];
    }
  }
  if (y == nullptr) return;
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    if (c < D) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    if (c < D) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
      const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
      const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (OUT_BF16) {
        uint2 o;
        o.x = pack_bf16(o0, o1);
        o.y = pack_bf16(o2, o3);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * D + c) = o;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * D + c) = make_float4(o0, o1, o2, o3);
      }
    }
  }
}

// G = 1 with narrow rows (C = 96 / 192: Swin stages 1 and 2): 32 / LPR rows per warp, every lane carries NV float4
// (a whole warp per 96-float row leaves a quarter of the lanes idle; these launches cover 802 816 rows at batch 256).
template <int LPR, int NV, bool OUT_BF16>
__global__ void __launch_bounds__(256) gather_ln_sub_kernel(const float* __restrict__ x, const int* __restrict__ idx,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            void* __restrict__ y, float* __restrict__ copy, long long rows_out,
                                                            int T_in, int T_out, float eps) {
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  constexpr int C = 4 * LPR * NV;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long row = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  const bool live = row < rows_out;
  const long long r = live ? row : 0;
  const long long img = r / T_out;
  const int src = __ldg(idx + static_cast<int>(r - img * T_out));
  const float* xr = x + (img * T_in + src) * static_cast<long long>(C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = *reinterpret_cast<const float4*>(xr + (i * LPR + l) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  if (live && copy != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(copy + row * C + (i * LPR + l) * 4) = v[i];
  }
  if (y == nullptr) return;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
  if (!live) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * LPR + l) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
    const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
    if (OUT_BF16) {
      uint2 o;
      o.x = pack_bf16(o0, o1);
      o.y = pack_bf16(o2, o3);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * C + c) = o;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * C + c) = make_float4(o0, o1, o2, o3);
    }
  }
}

template <int LPR, int NV>
int launch_gather_ln_sub(const float* x, const int* idx, const float* gamma, const float* beta, void* y, int y_dtype, float* copy,
                         long long rows_out, int T_in, int T_out, float eps, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const unsigned grid = static_cast<unsigned>((rows_out + 8 * RPW - 1) / (8 * RPW));
  const bool pdl = pdl_for_work(rows_out, 4 * LPR * NV);
  if (y_dtype == EVT_BF16)
    EVT_CUDA(launch_pdl(gather_ln_sub_kernel<LPR, NV, true>, dim3(grid), dim3(256), 0, st, pdl, x, idx, gamma, beta, y, copy, rows_out,
                        T_in, T_out, eps));
  else
    EVT_CUDA(launch_pdl(gather_ln_sub_kernel<LPR, NV, false>, dim3(grid), dim3(256), 0, st, pdl, x, idx, gamma, beta, y, copy, rows_out,
                        T_in, T_out, eps));
  EVT_LAUNCH_CHECK("gather_ln_sub_kernel");
  return EVT_OK;
}

template <int NV>
int launch_gather_ln(const float* x, const int* idx, const float* gamma, const float* beta, void* y, int y_dtype, float* copy,
                     long long rows_out, int T_in, int T_out, int G, int C, float eps, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>((rows_out + 7) / 8);
  const bool pdl = pdl_for_work(rows_out, static_cast<long long>(G) * C);
  if (y_dtype == EVT_BF16)
    EVT_CUDA(launch_pdl(gather_ln_kernel<NV, true>, dim3(grid), dim3(256), 0, st, pdl, x, idx, gamma, beta, y, copy, rows_out, T_in,
                        T_out, G, C, eps));
  else
    EVT_CUDA(launch_pdl(gather_ln_kernel<NV, false>, dim3(grid), dim3(256), 0, st, pdl, x, idx, gamma, beta, y, copy, rows_out, T_in,
                        T_out, G, C, eps));
  EVT_LAUNCH_CHECK("gather_ln_kernel");
  return EVT_OK;
}

// ---------------------------------------------------------------------------------------------------- window attention
constexpr int kWTok = 49;     // 7 x 7 window
constexpr int kWHd = 32;      // head size of every Swin variant (C / heads = 96 / 3 = ... = 32)
constexpr int kWRows = 64;    // query / key rows per matrix in shared memory (49 + zero padding)
constexpr int kWKeyCols = 56; // key columns of the score tile (7 n-tiles of 8); columns 49..55 carry -inf in the table
constexpr int kWWarps = 4;
constexpr int kWMatBytes = kWRows * kWHd * 2;  // 4 KB

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of (row, 16-byte chunk) in a [64 rows][64 bytes] matrix: chunks XOR-swizzled by (row / 2) % 4, so the
// eight 16-byte rows an ldmatrix phase reads fall into eight different bank groups
__device__ __forceinline__ uint32_t wofs(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

// qkv  bf16 [n_windows * 49, ldq], columns q | k | v (each heads * 32), head h at h * 32 inside each
// tab  f32 [n_tab, heads, 64, 56] = (relative position bias + shift mask) * log2(e); key columns >= 49 hold -inf
// ctx  bf16 [n_windows * 49, ldc]
__global__ void __launch_bounds__(kWWarps * 32) window_attention_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq,
                                                                        __nv_bfloat16* __restrict__ ctx, long long ldc,
                                                                        const float* __restrict__ tab, int n_tab, int heads,
                                                                        long long n_items, float scale_log2e) {
  extern __shared__ __align__(128) uint8_t wsm[];
  ptx::grid_dep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sQ = wsm + warp * 3 * kWMatBytes;
  uint8_t* sK = sQ + kWMatBytes;
  uint8_t* sV = sK + kWMatBytes;
  // zero the padding rows 49..63 once (cp.async never touches them): P = 0 there must meet finite V, and padded K
  // rows only feed columns the table masks
  for (int i = lane; i < 3 * (kWRows - kWTok) * 4; i += 32) {
    const int mat = i / ((kWRows - kWTok) * 4), rc = i % ((kWRows - kWTok) * 4);
    *reinterpret_cast<uint4*>(sQ + mat * kWMatBytes + wofs(kWTok + rc / 4, rc % 4)) = make_uint4(0, 0, 0, 0);
  }
  ptx::grid_dep_wait();
  const int C = heads * kWHd;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV);
  for (long long item = static_cast<long long>(blockIdx.x) * kWWarps + warp; item < n_items;
       item += static_cast<long long>(gridDim.x) * kWWarps) {
    const long long w = item / heads;
    const int h = static_cast<int>(item - w * heads);
    const long long row0 = w * kWTok;
    __syncwarp();  // every lane is done with the previous item's shared memory
    for (int c = lane; c < 3 * kWTok * 4; c += 32) {
      const int mat = c / (kWTok * 4), rc = c - mat * (kWTok * 4);
      const int row = rc >> 2, ch = rc & 3;
      const __nv_bfloat16* src = qkv + (row0 + row) * ldq + mat * C + h * kWHd + ch * 8;
      const uint32_t dst = aQ + mat * kWMatBytes + wofs(row, ch);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // K fragments (B operand of S = Q K^T): n-tile nt = keys 8 nt .. 8 nt + 7; registers {ks0.b0, ks0.b1, ks1.b0, ks1.b1}
    uint32_t kb[7][4];
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) ldsm_x4(aK + wofs(nt * 8 + (lane & 7), lane >> 3), kb[nt]);
    // V fragments (B operand of O = P V): k-step kk = keys 16 kk .. +15, dims chunk pair cp: {n 2cp: b0, b1, n 2cp+1: b0, b1}
    uint32_t vb[4][2][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int cp = 0; cp < 2; ++cp)
        ldsm_x4_t(aV + wofs(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, cp * 2 + (lane >> 4)), vb[kk][cp]);
    const float* tb = tab + ((w % n_tab) * heads + h) * static_cast<long long>(kWRows * kWKeyCols);
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t qa[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) ldsm_x4(aQ + wofs(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)), qa[ks]);
      float s[7][4];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        mma_bf16(s[nt], qa[0], kb[nt][0], kb[nt][1]);
        mma_bf16(s[nt], qa[1], kb[nt][2], kb[nt][3]);
      }
      // scores -> log2 domain with bias / mask; rows g and g + 8 of this m-tile, columns 8 nt + 2 t, + 1
      const float* t0 = tb + (mt * 16 + g) * kWKeyCols + 2 * t;
      const float* t1 = t0 + 8 * kWKeyCols;
      float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        const float2 b0 = __ldg(reinterpret_cast<const float2*>(t0 + nt * 8));
        const float2 b1 = __ldg(reinterpret_cast<const float2*>(t1 + nt * 8));
        s[nt][0] = fmaf(s[nt][0], scale_log2e, b0.x);
        s[nt][1] = fmaf(s[nt][1], scale_log2e, b0.y);
        s[nt][2] = fmaf(s[nt][2], scale_log2e, b1.x);
        s[nt][3] = fmaf(s[nt][3], scale_log2e, b1.y);
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
      uint32_t pa[4][4];  // probabilities as A fragments of the four k-steps (keys 56..63 do not exist -> 0)
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        const float p0 = ex2f(s[nt][0] - m0), p1 = ex2f(s[nt][1] - m0);
        const float p2 = ex2f(s[nt][2] - m1), p3 = ex2f(s[nt][3] - m1);
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      pa[3][2] = pa[3][3] = 0u;
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      float o[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16(o[n], pa[kk], vb[kk][n >> 1][(n & 1) * 2], vb[kk][n >> 1][(n & 1) * 2 + 1]);
      const float i0 = 1.0f / l0, i1 = 1.0f / l1;  // l >= 1: the row maximum contributes 2^0
      const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int col = h * kWHd + n * 8 + 2 * t;
        if (r0 < kWTok) *reinterpret_cast<uint32_t*>(ctx + (row0 + r0) * ldc + col) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
        if (r1 < kWTok) *reinterpret_cast<uint32_t*>(ctx + (row0 + r1) * ldc + col) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- window attention, tcgen05
// EXPERIMENTAL (built with EVT_EXPERIMENTAL=1, selected with EVT_SWIN_ATTN_TC=1): correct -- the Swin op and model parity
// tests pass with it -- but SLOWER than the warp-level kernel above on B200: Swin-T batch 256 37.2 k img/s (34.4 k with the
// table row prefetched ahead of the S wait: more spills) against 40.4 k.  Two CTAs per SM leave 168 registers per thread for a
// row-per-thread softmax that wants 56 scores + 56 table values + 32 packed probabilities live (100-330 bytes of spills), and
// each head is a serial QK -> softmax -> PV -> read-out chain with a barrier round trip between the steps; the mma.sync kernel
// keeps a whole (window, head) in one warp's registers with no hand-offs.  A 49 x 49 x 32 problem is too small for the
// tcgen05 hand-off latencies to amortise.
#ifdef EVT_EXPERIMENTAL
// Two 49-token windows share one 128-row tcgen05 tile (window 0 in TMEM lanes / key columns 0..48, window 1 in
// 64..112): S = Q K^T is ONE M = 128, N = 128 MMA pair per head whose off-diagonal 64 x 64 blocks are simply never read,
// P (bf16, written back into TMEM over the consumed scores) is block-diagonal by construction -- zeros outside the row's
// own window -- so O = P V over all 128 key rows is exact.  Two heads (2 x 32 columns = one 128-byte row) share the Q / K / V
// tiles; the head is selected by the K offset of the operand descriptors.
//   work item         (window pair, head pair); persistent CTAs, TWO per SM (256 TMEM columns and ~113 KB of shared memory
//                     each), so one CTA's serial chain QK -> softmax -> PV -> read-out overlaps the other's
//   warp 4 (1 thread) TMA producer: Q, K, V as two 64-row boxes each (one per window) into a 2-deep ring
//   warp 5 (1 thread) MMA issuer
//   warps 0-3         thread = query row: scores * scale + (relative position bias + shift mask) table, max, exp2, bf16 P,
//                     sum; then O / sum -> bf16 into a 128B-swizzled staging tile; one 49-row TMA store per window
constexpr int kTcStages = 2;
constexpr int kTcTile = 128 * 128;               // one operand tile: 128 rows x 64 bf16
constexpr int kTcStageBytes = 3 * kTcTile;       // Q | K | V
constexpr int kTcStgBytes = (56 + kWTok) * 128;  // staging: window 0 at row 0, window 1 at row 56 (1024-byte aligned)
constexpr int kTcThreads = 192;
constexpr int kTcSCols = 128, kTcOCol = 128, kTcTmemCols = 256;

struct WinTcParams {
  const float* tab;
  long long n_windows, n_items;
  int n_tab, heads, hp_count;
  float scale_log2e;
};

__global__ void __launch_bounds__(kTcThreads, 2)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, const WinTcParams p) {
  extern __shared__ uint8_t tc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kTcStages * kTcStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kTcStgBytes);
  uint64_t* full = bars;            // [2] producer -> issuer
  uint64_t* empty = bars + 2;       // [2] issuer (commit after the item's last PV) -> producer
  uint64_t* s_ready = bars + 4;     // issuer (commit) -> softmax warps
  uint64_t* p_ready = bars + 5;     // softmax warps (4 arrivals) -> issuer
  uint64_t* o_ready = bars + 6;     // issuer (commit) -> softmax warps
  uint64_t* slot_free = bars + 7;   // softmax warps (4 arrivals): O read out, the slot may take the next S
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.heads * kWHd;
  const long long my_items = (p.n_items - static_cast<long long>(blockIdx.x) + gridDim.x - 1) / gridDim.x;

  if (warp == 4 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmIn);
    ptx::prefetch_tmap(&tmOut);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(s_ready, 1);
    ptx::mbar_init(p_ready, 4);
    ptx::mbar_init(o_ready, 1);
    ptx::mbar_init(slot_free, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 5) ptx::tmem_alloc<kTcTmemCols>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();

  if (warp == 4) {
    if (ptx::elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (long long it = 0; it < my_items; ++it) {
        const long long item = blockIdx.x + it * gridDim.x;
        const long long wp = item / p.hp_count;
        const int hp = static_cast<int>(item - wp * p.hp_count);
        ptx::mbar_wait(&empty[st], ph ^ 1);
        uint8_t* base = smem + st * kTcStageBytes;
        ptx::mbar_arrive_expect_tx(&full[st], kTcStageBytes);
        for (int wi = 0; wi < 2; ++wi) {
          const long long row = (2 * wp + wi) * kWTok;   // a window past the end loads zeros (rows out of range)
          const int r = row < 0x7fffffffll ? static_cast<int>(row) : 0x7fffffff;
          for (int m = 0; m < 3; ++m)
            ptx::tma_load_2d(base + m * kTcTile + wi * (kTcTile / 2), &tmIn, &full[st], m * C + hp * 64, r);
        }
        if (++st == kTcStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_qk = ptx::make_idesc(128, 128, 1, 0, 0);
      constexpr uint32_t idesc_pv = ptx::make_idesc(128, 64, 1, 0, 1);
      int st = 0;
      uint32_t ph = 0, hph = 0;  // hph: phase of the per-head barriers (one completion per head processed)
      for (long long it = 0; it < my_items; ++it) {
        const long long item = blockIdx.x + it * gridDim.x;
        const int hp = static_cast<int>(item % p.hp_count);
        const int nh = p.heads - 2 * hp < 2 ? p.heads - 2 * hp : 2;
        ptx::mbar_wait(&full[st], ph);
        const uint32_t sq = ptx::smem_u32(smem + st * kTcStageBytes);
        const uint64_t qd = ptx::smem_desc_sw128(sq), kd = ptx::smem_desc_sw128(sq + kTcTile), vd = ptx::smem_desc_sw128(sq + 2 * kTcTile);
        for (int hl = 0; hl < nh; ++hl) {
          ptx::mbar_wait(slot_free, hph ^ 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 2; ++k)  // head hl = K offsets 32 hl .. 32 hl + 31 of the 64-column tiles
            ptx::mma_f16_ss(tmem_base, qd + 2 * (2 * hl + k), kd + 2 * (2 * hl + k), idesc_qk, k != 0 ? 1u : 0u);
          ptx::mma_commit(s_ready);
          ptx::mbar_wait(p_ready, hph);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 keys per step: 8 TMEM columns of P, 2048 bytes of V (MN-major)
            ptx::mma_f16_ts(tmem_base + kTcOCol, tmem_base + 8 * k, vd + static_cast<uint64_t>(128 * k), idesc_pv, k != 0 ? 1u : 0u);
          ptx::mma_commit(o_ready);
          if (hl == nh - 1) ptx::mma_commit(&empty[st]);
          hph ^= 1;
        }
        if (++st == kTcStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    const int r = threadIdx.x;                 // query row of the tile = TMEM lane
    const int wi = r >> 6, i = r & 63;         // window of the pair, token in the window
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* stg_row = staging + (wi * 56 + i) * 128;
    const int sw = (wi * 56 + i) & 7;
    uint32_t hph = 0;
    for (long long it = 0; it < my_items; ++it) {
      const long long item = blockIdx.x + it * gridDim.x;
      const long long wp = item / p.hp_count;
      const int hp = static_cast<int>(item - wp * p.hp_count);
      const int nh = p.heads - 2 * hp < 2 ? p.heads - 2 * hp : 2;
      const long long w = 2 * wp + wi;
      const bool live = i < kWTok && w < p.n_windows;
      // tcgen05.ld / .st are warp-collective (.sync.aligned): they are issued under the warp-uniform condition and the rows
      // past a window's 49 tokens (lanes 17..31 of warps 1 and 3) only discard what they loaded
      const bool warp_live = w < p.n_windows;
      for (int hl = 0; hl < nh; ++hl) {
        const int h = 2 * hp + hl;
        const float* tb = p.tab + (((w % p.n_tab) * p.heads + h) * static_cast<long long>(kWRows) + i) * kWKeyCols;
        // the row of the (bias + mask) table is requested BEFORE the wait for S = Q K^T: its L2 latency hides behind the MMA
        float t[kWKeyCols];
        if (live) {
#pragma unroll
          for (int j = 0; j < kWKeyCols; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(tb + j));
            t[j] = bv.x, t[j + 1] = bv.y, t[j + 2] = bv.z, t[j + 3] = bv.w;
          }
        }
        ptx::mbar_wait(s_ready, hph);
        ptx::tc_fence_after();
        float sum = 1.f;
        uint32_t pk[32];   // this row's 64 keys as bf16 pairs (keys >= 56 and masked keys: 0)
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = 0u;
        uint32_t sc[64];
        if (warp_live) {
          uint32_t a[32], b[32];
          ptx::tmem_ld_x32(t_row + wi * 64, a);
          ptx::tmem_ld_x32(t_row + wi * 64 + 32, b);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[j] = a[j], sc[32 + j] = b[j];
        }
        if (live) {
          float m = -CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < kWKeyCols; j += 4) {
            t[j] = fmaf(__uint_as_float(sc[j]), p.scale_log2e, t[j]);
            t[j + 1] = fmaf(__uint_as_float(sc[j + 1]), p.scale_log2e, t[j + 1]);
            t[j + 2] = fmaf(__uint_as_float(sc[j + 2]), p.scale_log2e, t[j + 2]);
            t[j + 3] = fmaf(__uint_as_float(sc[j + 3]), p.scale_log2e, t[j + 3]);
            m = fmaxf(fmaxf(m, fmaxf(t[j], t[j + 1])), fmaxf(t[j + 2], t[j + 3]));
          }
          float l = 0.f;
#pragma unroll
          for (int j = 0; j < kWKeyCols; j += 2) {
            const float p0 = ex2f(t[j] - m), p1 = ex2f(t[j + 1] - m);   // masked keys: 2^-inf = 0
            l += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
          sum = l;
        }
        // block-diagonal P: window wi's keys live in columns 32 wi .. 32 wi + 31 of the packed row, zeros elsewhere
        {
          uint32_t z[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] = 0u;
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) lo[j] = pk[j], hi[j] = pk[16 + j];
          if (wi == 0) {
            ptx::tmem_st_x16(t_row, lo);
            ptx::tmem_st_x16(t_row + 16, hi);
            ptx::tmem_st_x16(t_row + 32, z);
            ptx::tmem_st_x16(t_row + 48, z);
          } else {
            ptx::tmem_st_x16(t_row, z);
            ptx::tmem_st_x16(t_row + 16, z);
            ptx::tmem_st_x16(t_row + 32, lo);
            ptx::tmem_st_x16(t_row + 48, hi);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(p_ready);
        const float inv = 1.0f / sum;   // sum >= 1: the row maximum contributes 2^0
        ptx::mbar_wait(o_ready, hph);
        ptx::tc_fence_after();
        uint32_t o[32];
        if (warp_live) {
          ptx::tmem_ld_x32(t_row + kTcOCol + 32 * hl, o);   // O holds both heads' dims; this head's are columns 32 hl ..
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(slot_free);
        if (hl == 0 && it > 0) {  // the previous item's TMA stores have read the staging tile (long ago): safe to rewrite it
          if (warp == 0 && ptx::elect_one()) ptx::bulk_wait_read<0>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (live) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[8 * c4]) * inv, __uint_as_float(o[8 * c4 + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(o[8 * c4 + 2]) * inv, __uint_as_float(o[8 * c4 + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(o[8 * c4 + 4]) * inv, __uint_as_float(o[8 * c4 + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(o[8 * c4 + 6]) * inv, __uint_as_float(o[8 * c4 + 7]) * inv);
            *reinterpret_cast<uint4*>(stg_row + (((4 * hl + c4) ^ sw) << 4)) = v;
          }
        }
        hph ^= 1;
      }
      // both heads of the pair are staged: one 49-row store per window (columns past C are clipped by the tensor map)
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 0 && ptx::elect_one()) {
        const long long w0 = 2 * wp;
        ptx::tma_store_2d(&tmOut, staging, hp * 64, static_cast<int>(w0 * kWTok));
        if (w0 + 1 < p.n_windows) ptx::tma_store_2d(&tmOut, staging + 56 * 128, hp * 64, static_cast<int>((w0 + 1) * kWTok));
        ptx::bulk_commit();
      }
    }
    if (warp == 0 && ptx::elect_one()) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTcTmemCols>(tmem_base);
  }
}

#endif  // EVT_EXPERIMENTAL

// ---------------------------------------------------------------------------------------------------- final LN + pool
// One CTA (8 warps) per image: every warp normalises tokens warp, warp + 8, ... and accumulates them; the warps' sums meet
// in shared memory.  D % 4 == 0, D <= 128 * NV.
template <int NV>
__global__ void __launch_bounds__(256) ln_mean_tokens_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                             int T, int D, float eps) {
  __shared__ float acc_s[8][128 * NV];
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<long long>(blockIdx.x) * T * D;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tk = warp; tk < T; tk += 8) {
    const float* xr = xb + static_cast<long long>(tk) * D;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < D) {
        v[i] = *reinterpret_cast<const float4*>(xr + c);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) / static_cast<float>(D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < D) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      acc[i].x += (v[i].x - mean) * rstd;
      acc[i].y += (v[i].y - mean) * rstd;
      acc[i].z += (v[i].z - mean) * rstd;
      acc[i].w += (v[i].w - mean) * rstd;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(&acc_s[warp][i * 128 + lane * 4]) = acc[i];
  __syncthreads();
  // mean_t(LN(x_t)) = gamma * mean_t(xhat_t) + beta
  const float inv_t = 1.0f / static_cast<float>(T);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += acc_s[w8][c];
    y[static_cast<long long>(blockIdx.x) * D + c] = __float2bfloat16_rn(s * inv_t * gamma[c] + beta[c]);
  }
}

// ---------------------------------------------------------------------------------------------------- im2col, P % 4 == 0
// One thread = 4 consecutive pixels of one image row (16 B in, 8 B out); same (c, i, j) K order as im2col_kernel.
__global__ void __launch_bounds__(256) im2col4_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ cols, int H,
                                                      int W, int P, long long total) {
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= total) return;
  const int w4 = W / 4;
  const int x4 = static_cast<int>(tid % w4);
  long long r = tid / w4;
  const int y = static_cast<int>(r % H);
  r /= H;
  const int c = static_cast<int>(r % 3);
  const long long b = r / 3;
  const float4 v = *reinterpret_cast<const float4*>(px + tid * 4);
  const int x = x4 * 4;
  const int py = y / P, i = y % P, pxi = x / P, j = x % P;
  const long long row = (b * (H / P) + py) * (W / P) + pxi;
  const int k = (c * P + i) * P + j;
  uint2 o;
  o.x = pack_bf16(v.x, v.y);
  o.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(cols + row * (3ll * P * P) + k) = o;
}

}  // namespace

int im2col4_launch(const float* pixels, void* cols, int B, int H, int W, int P, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * 3 * H * (W / 4);
  EVT_CUDA(launch_pdl(im2col4_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, false, pixels,
                      reinterpret_cast<__nv_bfloat16*>(cols), H, W, P, total));
  EVT_LAUNCH_CHECK("im2col4_kernel");
  return EVT_OK;
}

}  // namespace evt

using namespace evt;

extern "C" int evt_gather_layernorm(const float* x, const int* idx, const float* gamma, const float* beta, void* y, int y_dtype,
                                    float* copy_f32, int64_t images, int T_in, int T_out, int G, int C, float eps,
                                    evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  EVT_CHECK_ARG(x && idx, "gather_layernorm: null pointer");
  EVT_CHECK_ARG(y != nullptr || copy_f32 != nullptr, "gather_layernorm: nothing to write");
  EVT_CHECK_ARG(y == nullptr || (gamma && beta), "gather_layernorm: LayerNorm output needs gamma and beta");
  EVT_CHECK_ARG(images > 0 && T_in > 0 && T_out > 0 && (G == 1 || G == 4), "gather_layernorm: bad sizes (G must be 1 or 4)");
  EVT_CHECK_ARG(C > 0 && C % 4 == 0 && G * C <= 3072, "gather_layernorm: C must be a multiple of 4 with G*C <= 3072");
  EVT_CHECK_ARG(y_dtype == EVT_BF16 || y_dtype == EVT_F32, "gather_layernorm: y dtype must be bf16 or f32");
  EVT_CHECK_ARG(eps >= 0.f, "gather_layernorm: negative eps");
  const long long rows = images * T_out;
  const int D = G * C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (G == 1 && C == 96) return launch_gather_ln_sub<8, 3>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, eps, st);
  if (G == 1 && C == 192) return launch_gather_ln_sub<16, 3>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, eps, st);
  const int nv = (D + 127) / 128;
  if (nv <= 1) return launch_gather_ln<1>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, G, C, eps, st);
  if (nv <= 2) return launch_gather_ln<2>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, G, C, eps, st);
  if (nv <= 3) return launch_gather_ln<3>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, G, C, eps, st);
  if (nv <= 6) return launch_gather_ln<6>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, G, C, eps, st);
  if (nv <= 12) return launch_gather_ln<12>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, G, C, eps, st);
  return launch_gather_ln<24>(x, idx, gamma, beta, y, y_dtype, copy_f32, rows, T_in, T_out, G, C, eps, st);
}

extern "C" int evt_window_attention_fwd(const void* qkv, int64_t ldq, void* ctx, int64_t ldc, const float* table, int n_tab,
                                        int64_t n_windows, int window_tokens, int heads, int head_size, float scale,
                                        evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  EVT_CHECK_ARG(qkv && ctx && table, "window_attention: null pointer");
  EVT_CHECK_ARG(n_windows > 0 && heads > 0 && n_tab > 0, "window_attention: sizes must be positive");
  if (window_tokens != kWTok || head_size != kWHd)
    return fail(EVT_ERR_UNSUPPORTED, "window_attention: only 7x7 windows (49 tokens) with head size 32 are implemented");
  EVT_CHECK_ARG(ldq >= 3ll * heads * kWHd && ldc >= static_cast<int64_t>(heads) * kWHd, "window_attention: leading dimension too small");
  EVT_CHECK_ARG(ldq % 8 == 0 && reinterpret_cast<uintptr_t>(qkv) % 16 == 0, "window_attention: qkv rows must be 16-byte aligned");
  EVT_CHECK_ARG(ldc % 2 == 0 && reinterpret_cast<uintptr_t>(ctx) % 4 == 0, "window_attention: ctx rows must be 4-byte aligned");
#ifdef EVT_EXPERIMENTAL
  static const bool tc_on = getenv("EVT_SWIN_ATTN_TC") != nullptr && atoi(getenv("EVT_SWIN_ATTN_TC")) != 0;  // A/B timing
  if (tc_on && ldc % 8 == 0 && reinterpret_cast<uintptr_t>(ctx) % 16 == 0 && n_windows * kWTok < (1ll << 31) - 256) {
    const int C = heads * kWHd;
    CUtensorMap tmIn, tmOut;
    rc = make_tmap_2d(&tmIn, qkv, 2, static_cast<uint64_t>(n_windows) * kWTok, 3ull * C, static_cast<uint64_t>(ldq), 64, 64);
    if (rc != EVT_OK) return rc;
    rc = make_tmap_2d(&tmOut, ctx, 2, static_cast<uint64_t>(n_windows) * kWTok, static_cast<uint64_t>(C), static_cast<uint64_t>(ldc), kWTok, 64);
    if (rc != EVT_OK) return rc;
    WinTcParams q;
    q.tab = table;
    q.n_windows = n_windows;
    q.n_tab = n_tab;
    q.heads = heads;
    q.hp_count = (heads + 1) / 2;
    q.n_items = ((n_windows + 1) / 2) * q.hp_count;
    q.scale_log2e = scale * 1.4426950408889634f;
    const int smem_tc = 1024 + kTcStages * kTcStageBytes + kTcStgBytes + 128;
    static int tc_dev = -1;
    int dev_tc = 0;
    EVT_CUDA(cudaGetDevice(&dev_tc));
    if (tc_dev != dev_tc) {
      EVT_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc));
      tc_dev = dev_tc;
    }
    const long long cap_tc = 2ll * num_sms();
    const unsigned grid_tc = static_cast<unsigned>(q.n_items < cap_tc ? q.n_items : cap_tc);
    EVT_CUDA(launch_pdl(window_attention_tc_kernel, dim3(grid_tc), dim3(kTcThreads), smem_tc, static_cast<cudaStream_t>(stream),
                        pdl_for_work(n_windows * kWTok, static_cast<long long>(heads) * kWHd), tmIn, tmOut, q));
    EVT_LAUNCH_CHECK("window_attention_tc_kernel");
    return EVT_OK;
  }
#endif  // EVT_EXPERIMENTAL
  const long long items = n_windows * heads;
  const int smem = kWWarps * 3 * kWMatBytes;
  static int configured_dev = -1;
  int dev = 0;
  EVT_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    EVT_CUDA(cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_dev = dev;
  }
  const long long ctas_needed = (items + kWWarps - 1) / kWWarps;
  const long long cap = static_cast<long long>(num_sms()) * 4;  // 4 CTAs of 48 KB per SM, grid-stride over the rest
  const unsigned grid = static_cast<unsigned>(ctas_needed < cap ? ctas_needed : cap);
  EVT_CUDA(launch_pdl(window_attention_kernel, dim3(grid), dim3(kWWarps * 32), smem, static_cast<cudaStream_t>(stream),
                      pdl_for_work(n_windows * kWTok, static_cast<long long>(heads) * kWHd), reinterpret_cast<const __nv_bfloat16*>(qkv), static_cast<long long>(ldq),
                      reinterpret_cast<__nv_bfloat16*>(ctx), static_cast<long long>(ldc), table, n_tab, heads, items,
                      scale * 1.4426950408889634f));
  EVT_LAUNCH_CHECK("window_attention_kernel");
  return EVT_OK;
}

extern "C" int evt_layernorm_mean_tokens(const float* x, const float* gamma, const float* beta, void* y, int64_t images, int T,
                                         int D, float eps, evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  EVT_CHECK_ARG(x && gamma && beta && y, "layernorm_mean_tokens: null pointer");
  EVT_CHECK_ARG(images > 0 && T > 0 && D > 0 && D % 4 == 0 && D <= 1536, "layernorm_mean_tokens: D must be a multiple of 4, <= 1536");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(static_cast<unsigned>(images));
  __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y);
  const int nv = (D + 127) / 128;
  if (nv <= 3) EVT_CUDA(launch_pdl(ln_mean_tokens_kernel<3>, grid, dim3(256), 0, st, false, x, gamma, beta, yo, T, D, eps));
  else if (nv <= 6) EVT_CUDA(launch_pdl(ln_mean_tokens_kernel<6>, grid, dim3(256), 0, st, false, x, gamma, beta, yo, T, D, eps));
  else EVT_CUDA(launch_pdl(ln_mean_tokens_kernel<12>, grid, dim3(256), 0, st, false, x, gamma, beta, yo, T, D, eps));
  EVT_LAUNCH_CHECK("ln_mean_tokens_kernel");
  return EVT_OK;
}
