// out[M,N] = epilogue(A[M,K] . W[N,K]^T): persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 8      TMA producer   : A tile 128x64 and W tile BNx64 (bf16; 128x32 / BNx32 for tf32) into a
//                                 STAGES-deep ring of 128B-swizzled shared-memory buffers
//   warp 9      MMA issuer     : one thread issues tcgen05.mma (M=128, N=BN, K=16|8) into one of two
//                                 TMEM accumulator stages (2 x BN columns), commits to mbarriers
//   warps 0-7   epilogue       : tcgen05.ld the accumulator (lane == row; two warps per TMEM lane quadrant, taking
//                                 alternate 128-byte column chunks), bias / GELU on packed f32x2 pairs, then through
//                                 a 128B-swizzled shared tile to a TMA store / TMA f32 reduce-add (or coalesced
//                                 global stores on the generic path); overlaps the next tile's MMAs
//
// Tiles are walked N-fastest so the CTAs resident at one time share a few A row-blocks (read from HBM
// once) while the whole weight matrix stays in L2.  M, N and K tails are handled by TMA zero fill on the
// load side and by clipping / predication on the store side, so no operand is ever padded in HBM (only leading
// dimensions must be multiples of 16 bytes).
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace evt {
namespace {

constexpr int BM = 128;
constexpr int kStageRowBytes = 128;  // one swizzle row: 64 bf16 or 32 tf32 of K
constexpr int kEpiWarps = 8;
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = 32 * (kEpiWarps + 2);
constexpr int kStgBytes = 32 * 128;  // per epilogue warp: 32 dense 128-byte rows, 128B-swizzled

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * kStageRowBytes;
  static constexpr int kBBytes = BN * kStageRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 4 : BN == 192 ? 4 : BN == 128 ? 6 : 8;
  static constexpr int kTmemCols = BN == 256 ? 512 : BN == 192 ? 512 : BN == 128 ? 256 : 128;
  static constexpr int kStagingBytes = kEpiWarps * kStgBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kStagingBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

struct GemmParams {
  const float* bias;
  const float* residual;
  void* out;
  long long ldr, ldo;
  int M, N, K;
  int res_row_mod, res_row_off;
  int out_group, out_group_stride, out_group_off;
  int tiles_m, tiles_n, num_kb;
  int k_step;  // elements of K per stage (64 bf16 / 32 tf32)
  int vec_ok;  // out / residual / bias allow 16-byte vector access
  int round_tf32;  // f32 output feeds a tf32 tensor-core op: round to nearest tf32 when written
};

__device__ __forceinline__ long long map_out_row(const GemmParams& p, long long row) {
  if (p.out_group <= 0) return row;
  const long long g = row / p.out_group;
  return g * p.out_group_stride + p.out_group_off + (row - g * p.out_group);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed f32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 work on an aligned register pair) ----
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// erf-GELU of a pair, in place.  erfc(z) = exp2(z Q(z)) on z = |x|/sqrt(2) in [0, 4] with Q a degree-6 minimax fit
// of log2(erfc(z))/z (tools/fit_erfc.py: |gelu error| < 7e-7 absolute, < 7e-6 relative; erfc(4) = 1.5e-8 is below
// f32 resolution of 1, so z is clamped there).  The polynomial is written in a = |x| (powers of 1/sqrt(2) folded
// into the coefficients).  gelu(x) = x/2 (1 + sign(x)(1 - erfc)) = (x/2 + |x/2|) - |x/2| erfc.
// Per element: 1 MUFU (ex2), 3.5 packed FMA-class ops, 3 scalar ops -- the A&S 7.1.26 form needs 2 MUFU + 13 scalar.
__device__ __forceinline__ void gelu_erf_pair(float& x0, float& x1) {
  constexpr float kAmax = 5.65685424949238f;  // 4 sqrt(2)
  const uint64_t a = pk2(fminf(fabsf(x0), kAmax), fminf(fabsf(x1), kAmax));
  uint64_t q = fma2(a, pk2(-1.765649017e-06f, -1.765649017e-06f), pk2(6.025074981e-05f, 6.025074981e-05f));
  q = fma2(q, a, pk2(-9.201008943e-04f, -9.201008943e-04f));
  q = fma2(q, a, pk2(8.467212319e-03f, 8.467212319e-03f));
  q = fma2(q, a, pk2(-5.387612060e-02f, -5.387612060e-02f));
  q = fma2(q, a, pk2(-4.585517347e-01f, -4.585517347e-01f));
  q = fma2(q, a, pk2(-1.151212096e+00f, -1.151212096e+00f));
  float t0, t1;
  upk2(mul2(q, a), t0, t1);
  const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
  float h0, h1;
  upk2(mul2(pk2(x0, x1), pk2(0.5f, 0.5f)), h0, h1);
  x0 = fmaf(-fabsf(h0), e0, h0 + fabsf(h0));
  x1 = fmaf(-fabsf(h1), e1, h1 + fabsf(h1));
}
// tanh-GELU of a pair: 0.5 x (1 + tanh(u)) == x * sigmoid(2u) ; u = sqrt(2/pi) (x + 0.044715 x^3)
__device__ __forceinline__ void gelu_tanh_pair(float& x0, float& x1) {
  constexpr float k0 = -2.0f * 1.4426950408889634f * 0.7978845608028654f;  // exponent of 2 is -2 log2(e) u
  constexpr float k1 = k0 * 0.044715f;
  const uint64_t x = pk2(x0, x1);
  const uint64_t w = fma2(mul2(x, x), pk2(k1, k1), pk2(k0, k0));
  float t0, t1;
  upk2(mul2(x, w), t0, t1);
  x0 *= rcp_approx(1.0f + ex2_approx(t0));
  x1 *= rcp_approx(1.0f + ex2_approx(t1));
}
template <int ACT, bool EXACT>
__device__ __forceinline__ void apply_act_pair(float& x0, float& x1) {
  if (ACT == EVT_ACT_GELU_ERF) {
    if (EXACT) {
      x0 = 0.5f * x0 * (1.0f + erff(x0 * 0.70710678118654752f));
      x1 = 0.5f * x1 * (1.0f + erff(x1 * 0.70710678118654752f));
    } else {
      gelu_erf_pair(x0, x1);
    }
  } else if (ACT == EVT_ACT_GELU_TANH) {
    if (EXACT) {
      x0 = 0.5f * x0 * (1.0f + tanhf(0.7978845608028654f * (x0 + 0.044715f * x0 * x0 * x0)));
      x1 = 0.5f * x1 * (1.0f + tanhf(0.7978845608028654f * (x1 + 0.044715f * x1 * x1 * x1)));
    } else {
      gelu_tanh_pair(x0, x1);
    }
  }
}
// v[j] = act(v[j] + bias[n0 + j]) for the CH columns of one chunk (columns >= N are don't-care).
template <int CH, int ACT, bool EXACT>
__device__ __forceinline__ void bias_act(float (&v)[CH], const float* __restrict__ bias, int n0, int N, bool full) {
  if (bias != nullptr) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
      for (int i = 0; i < CH / 4; ++i) {
        const float4 b = __ldg(b4 + i);
        upk2(add2(pk2(v[4 * i], v[4 * i + 1]), pk2(b.x, b.y)), v[4 * i], v[4 * i + 1]);
        upk2(add2(pk2(v[4 * i + 2], v[4 * i + 3]), pk2(b.z, b.w)), v[4 * i + 2], v[4 * i + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (n0 + j < N) v[j] += __ldg(bias + n0 + j);
    }
  }
  if (ACT != EVT_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < CH; j += 2) apply_act_pair<ACT, EXACT>(v[j], v[j + 1]);
  }
}

template <int BN, bool TF32, bool OUT_F32, int ACT, bool TMA_OUT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + C::kStages * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == kProducerWarp && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    if (TMA_OUT) ptx::prefetch_tmap(&tmO);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) ptx::tmem_alloc<C::kTmemCols>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          ptx::mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
          ptx::tma_load_2d(sa, &tmA, &full[stage], kb * p.k_step, m0);
          ptx::tma_load_2d_hint(sb, &tmW, &full[stage], kb * p.k_step, n0, ptx::kEvictLast);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(BM, BN, TF32 ? 2 : 1, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(stage_base + stage * C::kStageBytes);
          const uint64_t adesc = ptx::smem_desc_sw128(sa);
          const uint64_t bdesc = ptx::smem_desc_sw128(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per stage row
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            if (TF32) ptx::mma_tf32_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
            else ptx::mma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
          }
          ptx::mma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::mma_commit(&tfull[as]);  // accumulator complete
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 0..7
    // Warp w reads TMEM lanes 32*(w%4).. (the hardware's lane quadrant rule) and takes the 128-byte output
    // chunks (64 bf16 / 32 f32 columns) c = w/4, w/4 + 2, ... of the tile, so two warps share every row block.
    // Thread == accumulator row while the bias / activation math runs on packed f32x2 pairs (FFMA2: half the
    // issue slots of scalar FMAs); the converted chunk then goes through a 128B-swizzled 4 KB smem tile.
    constexpr int CH = OUT_F32 ? 32 : 64;
    constexpr int NCH = BN / CH;
    const int quad = warp & 3;
    const int grp = warp >> 2;
    uint8_t* stg = staging + warp * kStgBytes;
    const int sw = lane & 7;
    int as = 0;
    uint32_t aphase = 0;
    const bool has_res = p.residual != nullptr;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.tiles_n) * BM + quad * 32;
      const int nt0 = (tile % p.tiles_n) * BN;
      const int rows_here = min(32, p.M - m0);  // may be <= 0 for a fully out-of-range warp
      // Residual rows of this warp's 32x32 chunk, one float4 per (4-row group, lane): issued one chunk ahead so
      // the HBM latency of the skip connection hides behind the previous chunk (and behind the MMA wait).
      float4 rnext[8];
      auto load_res = [&](int n0c) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3);
          rnext[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr < rows_here && p.vec_ok && n0c + CH <= p.N) {
            const long long row = m0 + rr;
            const long long rrow = p.res_row_mod > 0 ? p.res_row_off + row % p.res_row_mod : map_out_row(p, row);
            rnext[it] = *reinterpret_cast<const float4*>(p.residual + rrow * p.ldr + n0c + (lane & 7) * 4);
          }
        }
      };
      if constexpr (OUT_F32 && !TMA_OUT) {
        if (has_res && grp < NCH) load_res(nt0 + grp * CH);
      }
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      if (rows_here > 0) {
#pragma unroll 1
        for (int c = grp; c < NCH; c += 2) {
          const int n0 = nt0 + c * CH;
          if (n0 >= p.N) break;
          float v[CH];
          {
            uint32_t r[32];
            ptx::tmem_ld_x32(t_row + c * CH, r);
            if constexpr (CH == 64) {
              uint32_t r2[32];
              ptx::tmem_ld_x32(t_row + c * CH + 32, r2);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[32 + j] = __uint_as_float(r2[j]);
            } else {
              ptx::tmem_ld_wait();
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          const bool full = p.vec_ok && (n0 + CH <= p.N);
          float4 rcur[8];
          if constexpr (OUT_F32 && !TMA_OUT) {
            if (has_res) {
#pragma unroll
              for (int it = 0; it < 8; ++it) rcur[it] = rnext[it];
              if (c + 2 < NCH && n0 + 2 * CH < p.N) load_res(n0 + 2 * CH);
            }
          }
          bias_act<CH, ACT, TF32>(v, p.bias, n0, p.N, full);
          if constexpr (OUT_F32) {
            if (p.round_tf32) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = ptx::round_tf32(v[j]);
            }
          }
          if constexpr (!OUT_F32) if (has_res && lane < rows_here) {  // rare combination: add before rounding to bf16
            const long long orow = map_out_row(p, m0 + lane);
            const long long rrow = p.res_row_mod > 0 ? p.res_row_off + (m0 + lane) % p.res_row_mod : orow;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (n0 + j < p.N) v[j] += p.residual[rrow * p.ldr + n0 + j];
          }
          // stage: dense 128-byte rows, 16-byte pieces XOR-swizzled by (row & 7) -- the layout a SWIZZLE_128B
          // tensor map expects, and conflict-free for both the row-per-thread writes and the row-segment reads.
          if constexpr (TMA_OUT) {
            if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store of this warp has finished reading stg
          }
          __syncwarp();
          uint8_t* sb = stg + lane * 128;
          if constexpr (OUT_F32) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(sb + ((i ^ sw) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 w;
              w.x = pack_bf16x2(v[8 * i], v[8 * i + 1]);
              w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
              w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
              w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
              *reinterpret_cast<uint4*>(sb + ((i ^ sw) << 4)) = w;
            }
          }
          if constexpr (TMA_OUT) {
            // The TMA engine does the coalescing, the M/N clipping and (for the skip connection) the f32 add into
            // the residual stream at L2 -- the SM never reads the residual.
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (has_res) ptx::tma_reduce_add_2d(&tmO, stg, n0, m0);
              else ptx::tma_store_2d(&tmO, stg, n0, m0);
              ptx::bulk_commit();
            }
            continue;
          }
          __syncwarp();
          // store: lanes 8k..8k+7 cover one 128-byte row segment; 4 rows per instruction
          const int piece = lane & 7;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3);
            if (rr < rows_here) {
              const long long row = m0 + rr;
              const long long orow = map_out_row(p, row);
              const uint8_t* src = stg + rr * 128 + ((piece ^ (rr & 7)) << 4);
              if constexpr (OUT_F32) {
                float4 val = *reinterpret_cast<const float4*>(src);
                const int col = n0 + piece * 4;
                float* dst = reinterpret_cast<float*>(p.out) + orow * p.ldo + col;
                if (full) {
                  if (has_res) {
                    const float4 rs = rcur[it];
                    val.x += rs.x;
                    val.y += rs.y;
                    val.z += rs.z;
                    val.w += rs.w;
                  }
                  *reinterpret_cast<float4*>(dst) = val;
                } else {
                  const float e[4] = {val.x, val.y, val.z, val.w};
                  const long long rrow = p.res_row_mod > 0 ? p.res_row_off + row % p.res_row_mod : orow;
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    if (col + q < p.N) dst[q] = e[q] + (has_res ? p.residual[rrow * p.ldr + col + q] : 0.f);
                }
              } else {
                const uint4 val = *reinterpret_cast<const uint4*>(src);
                const int col = n0 + piece * 8;
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col;
                if (full) {
                  *reinterpret_cast<uint4*>(dst) = val;
                } else {
                  const uint32_t e[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                  for (int q = 0; q < 8; ++q)
                    if (col + q < p.N) {
                      const uint16_t h = static_cast<uint16_t>(e[q >> 1] >> ((q & 1) * 16));
                      reinterpret_cast<uint16_t*>(dst)[q] = h;
                    }
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (TMA_OUT && lane == 0) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

int choose_bn(int N) {
  const int cands[4] = {256, 192, 128, 64};
  int best = 256;
  long best_cost = -1;
  for (int bn : cands) {
    long cost = static_cast<long>((N + bn - 1) / bn) * bn;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

template <int BN, bool TF32, bool OUT_F32, int ACT, bool TMA_OUT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_kernel<BN, TF32, OUT_F32, ACT, TMA_OUT>;
  static bool configured = false;  // per instantiation; attribute is per-device-context but identical everywhere
  static int configured_dev = -1;
  int dev = 0;
  EVT_CUDA(cudaGetDevice(&dev));
  if (!configured || configured_dev != dev) {
    EVT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
    configured_dev = dev;
  }
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kern<<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmW, tmO, p);
  EVT_LAUNCH_CHECK("gemm_kernel");
  return EVT_OK;
}

template <int BN, bool TF32, bool TMA_OUT>
int dispatch_epi(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& o, const GemmParams& p, bool out_f32, int act,
                 cudaStream_t s) {
  if (out_f32) {
    switch (act) {
      case EVT_ACT_NONE: return launch<BN, TF32, true, EVT_ACT_NONE, TMA_OUT>(a, w, o, p, s);
      case EVT_ACT_GELU_ERF: return launch<BN, TF32, true, EVT_ACT_GELU_ERF, TMA_OUT>(a, w, o, p, s);
      default: return launch<BN, TF32, true, EVT_ACT_GELU_TANH, TMA_OUT>(a, w, o, p, s);
    }
  }
  switch (act) {
    case EVT_ACT_NONE: return launch<BN, TF32, false, EVT_ACT_NONE, TMA_OUT>(a, w, o, p, s);
    case EVT_ACT_GELU_ERF: return launch<BN, TF32, false, EVT_ACT_GELU_ERF, TMA_OUT>(a, w, o, p, s);
    default: return launch<BN, TF32, false, EVT_ACT_GELU_TANH, TMA_OUT>(a, w, o, p, s);
  }
}

}  // namespace

// in_dtype: EVT_BF16 (kind::f16) or EVT_F32 (kind::tf32, operands read as fp32 and truncated to tf32 by the MMA)
int gemm_launch(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, const float* bias,
                const float* residual, int64_t ldr, int res_row_mod, int res_row_off, void* out, int out_dtype,
                int64_t ldo, int out_group, int out_group_stride, int out_group_off, int64_t M, int N, int K, int act,
                cudaStream_t stream) {
  EVT_CHECK_ARG(A && W && out, "gemm: null pointer");
  EVT_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: M, N, K must be positive");
  EVT_CHECK_ARG(M < (1ll << 31) - 256, "gemm: M too large");
  EVT_CHECK_ARG(in_dtype == EVT_BF16 || in_dtype == EVT_F32, "gemm: input dtype must be bf16 or f32(tf32)");
  EVT_CHECK_ARG(out_dtype == EVT_BF16 || out_dtype == EVT_F32 || out_dtype == EVT_TF32, "gemm: out dtype must be bf16, f32 or tf32");
  EVT_CHECK_ARG(!(out_dtype == EVT_TF32 && residual != nullptr), "gemm: a tf32-rounded output cannot take a residual");
  EVT_CHECK_ARG(act >= EVT_ACT_NONE && act <= EVT_ACT_GELU_TANH, "gemm: unknown activation");
  EVT_CHECK_ARG(lda >= K && ldw >= K && ldo >= N, "gemm: leading dimension smaller than the row length");
  EVT_CHECK_ARG(out_group >= 0 && res_row_mod >= 0, "gemm: negative row-group parameter");
  EVT_CHECK_ARG(residual == nullptr || ldr >= N, "gemm: residual leading dimension smaller than N");
  const bool tf32 = in_dtype == EVT_F32;
  const int eb = tf32 ? 4 : 2;
  const int k_step = kStageRowBytes / eb;
  const int bn = choose_bn(N);
  CUtensorMap tmA, tmW;
  int rc = make_tmap_2d(&tmA, A, eb, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), BM, k_step);
  if (rc != EVT_OK) return rc;
  rc = make_tmap_2d(&tmW, W, eb, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw), bn, k_step);
  if (rc != EVT_OK) return rc;
  GemmParams p;
  p.bias = bias;
  p.residual = residual;
  p.out = out;
  p.ldr = ldr;
  p.ldo = ldo;
  p.M = static_cast<int>(M);
  p.N = N;
  p.K = K;
  p.res_row_mod = res_row_mod;
  p.res_row_off = res_row_off;
  p.out_group = out_group;
  p.out_group_stride = out_group_stride;
  p.out_group_off = out_group_off;
  p.tiles_m = static_cast<int>((M + BM - 1) / BM);
  p.tiles_n = (N + bn - 1) / bn;
  p.num_kb = (K + k_step - 1) / k_step;
  p.k_step = k_step;
  {
    const int oe = out_dtype != EVT_BF16 ? 4 : 2;
    bool ok = reinterpret_cast<uintptr_t>(out) % 16 == 0 && (ldo * oe) % 16 == 0;
    if (bias) ok = ok && reinterpret_cast<uintptr_t>(bias) % 16 == 0;
    if (residual) ok = ok && reinterpret_cast<uintptr_t>(residual) % 16 == 0 && (ldr * 4) % 16 == 0;
    p.vec_ok = ok ? 1 : 0;
  }
  const bool of32 = out_dtype != EVT_BF16;
  p.round_tf32 = out_dtype == EVT_TF32 ? 1 : 0;
  // TMA epilogue: contiguous output rows, 16-byte aligned, and the skip connection (if any) updated in place
  const int oeb = of32 ? 4 : 2;
  const bool tma_out = out_group == 0 && res_row_mod == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                       (ldo * oeb) % 16 == 0 &&
                       (residual == nullptr || (of32 && residual == reinterpret_cast<const float*>(out) && ldr == ldo));
  CUtensorMap tmO = tmA;
  if (tma_out) {
    rc = make_tmap_2d(&tmO, out, oeb, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldo), 32,
                      128 / oeb);
    if (rc != EVT_OK) return rc;
  }
#define EVT_BN_CASE(BNV)                                                                                   \
  case BNV:                                                                                                \
    if (tma_out)                                                                                           \
      return tf32 ? dispatch_epi<BNV, true, true>(tmA, tmW, tmO, p, of32, act, stream)                     \
                  : dispatch_epi<BNV, false, true>(tmA, tmW, tmO, p, of32, act, stream);                   \
    return tf32 ? dispatch_epi<BNV, true, false>(tmA, tmW, tmO, p, of32, act, stream)                      \
                : dispatch_epi<BNV, false, false>(tmA, tmW, tmO, p, of32, act, stream);
  switch (bn) {
    EVT_BN_CASE(256)
    EVT_BN_CASE(192)
    EVT_BN_CASE(128)
    EVT_BN_CASE(64)
  }
#undef EVT_BN_CASE
  return fail(EVT_ERR_INVALID, "gemm: no tile configuration");
}

}  // namespace evt

extern "C" int evt_gemm_bias_act(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                 const float* residual, int64_t ldr, int res_row_mod, int res_row_off, void* out,
                                 int out_dtype, int64_t ldo, int out_group, int out_group_stride, int out_group_off,
                                 int64_t M, int N, int K, int act, evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  return evt::gemm_launch(A, lda, W, ldw, EVT_BF16, bias, residual, ldr, res_row_mod, res_row_off, out, out_dtype, ldo,
                          out_group, out_group_stride, out_group_off, M, N, K, act, static_cast<cudaStream_t>(stream));
}

// tf32 flavour: A and W are f32 (the MMA reads them as tf32); out is f32.
extern "C" int evt_gemm_bias_act_tf32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                                      const float* residual, int64_t ldr, int res_row_mod, int res_row_off, float* out,
                                      int64_t ldo, int out_group, int out_group_stride, int out_group_off, int64_t M,
                                      int N, int K, int act, evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  return evt::gemm_launch(A, lda, W, ldw, EVT_F32, bias, residual, ldr, res_row_mod, res_row_off, out, EVT_F32, ldo,
                          out_group, out_group_stride, out_group_off, M, N, K, act, static_cast<cudaStream_t>(stream));
}
