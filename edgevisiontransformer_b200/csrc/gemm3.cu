// Residual GEMM with the NEXT LayerNorm fused into its epilogue (bf16 operands, CTA pairs, tcgen05 cta_group::2):
//
//     x[M,N]  <-  x + A[M,K] . W[N,K]^T + bias          (f32 residual stream, updated in place)
//     xn[M,N] <-  LayerNorm(x) * gamma + beta           (bf16, the A operand of the following QKV / FC1 GEMM)
//
// replaces  ViTSelfOutput.dense + residual + layernorm_after  and  ViTOutput.dense + residual + the next layer's
// layernorm_before  (SITE/models/vit/modeling_vit.py:265-268,308-312,333-340) -- one kernel instead of a TMA
// reduce-add GEMM followed by a LayerNorm kernel that re-reads the whole residual stream from HBM.
//
// A CTA pair owns a block of 256 rows and walks ALL N tiles of it back to back, so its epilogue warps see every
// column of their rows:
//   pass 1 (per 32-column chunk)  accumulator (TMEM, thread == row) -> 128B-swizzled smem tile -> coalesced layout
//                                 (8 lanes per row, 4 rows per instruction) + bias + old residual (prefetched one
//                                 chunk ahead with plain coalesced loads) -> new residual stored; per-lane shifted
//                                 sums (n, sum(x-K), sum((x-K)^2)) for the 8 rows a lane touches;
//   row statistics                shifted sums -> (mean, M2) per lane, Chan's pairwise combination over the 8 lanes
//                                 of a row (shuffles) and over the two warps sharing the rows (smem + named barrier):
//                                 centred variance, exact for constant rows even with eps = 1e-12;
//   pass 2                        every lane re-reads exactly the addresses it stored (L2 hits), normalises, writes
//                                 bf16 xn.  The MMAs of the next row block run underneath (two TMEM accumulators).
// Producer / MMA-issuer warps and the barrier protocol are those of gemm2.cu.
//
// STATUS: correct (tests/test_gpu_ops.py: residual bit-identical to the unfused path) but not used by the model runtime
// by default: on B200 the pass-2 re-read and the repeated A row-block reads miss L2 (ncu: 750 MB DRAM reads vs 465 MB
// algorithmic), and the eight epilogue warps cannot keep enough loads in flight -- 0.42 ms vs 0.22 ms unfused at 100k rows.
#include <stdio.h>
#include <stdlib.h>

#include "gemm_common.cuh"
#include "ops.h"

namespace evt {
namespace {

using namespace gemm_detail;

struct LnParams {
  const float* bias;   // [N] or null
  float* resid;        // [M, ldr] in/out
  const float* gamma;  // [N]
  const float* beta;   // [N]
  __nv_bfloat16* xn;   // [M, ldxn] out
  long long ldr, ldxn;
  int M, N, K;
  int tiles_n, num_kb, row_blocks;
  float eps, inv_n;
};

template <int BN>
struct Cfg3 {
  static constexpr int kABytes = BM * kStageRowBytes;
  static constexpr int kBBytes = (BN / 2) * kStageRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 6 : BN == 192 ? 6 : 8;
  static constexpr int kTmemCols = BN == 128 ? 256 : 512;
  static constexpr int kStagingBytes = kEpiWarps * kStgBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kStagingBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Chan et al. combination of two (count, mean, M2) summaries.
__device__ __forceinline__ void chan_combine(float& na, float& ma, float& qa, float nb, float mb, float qb) {
  const float n = na + nb;
  const float d = mb - ma;
  const float f = nb / n;
  ma = fmaf(d, f, ma);
  qa = qa + qb + d * d * na * f;
  na = n;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_pair_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const LnParams p) {
  using C = Cfg3<BN>;
  constexpr int CH = 32;          // f32 columns per chunk (128 bytes)
  constexpr int NCH = BN / CH;
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
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == kProducerWarp && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], 2 * kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) ptx::tmem_alloc_pair<C::kTmemCols>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::grid_dep_launch();  // the next kernel may begin its own prologue
  ptx::grid_dep_wait();    // operands / outputs of the previous kernel are complete from here on

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int rb = pair; rb < p.row_blocks; rb += num_pairs) {
        const int m0 = rb * (2 * BM) + static_cast<int>(rank) * BM;
        for (int nt = 0; nt < p.tiles_n; ++nt) {
          const int n0 = nt * BN + static_cast<int>(rank) * (BN / 2);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * C::kStageBytes;
            uint8_t* sb = sa + C::kABytes;
            const uint32_t lead_full = ptx::mapa(&full[stage], 0);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
            ptx::tma_load_2d_pair(sa, &tmA, lead_full, kb * 64, m0, ptx::kEvictFirst);
            ptx::tma_load_2d_pair(sb, &tmW, lead_full, kb * 64, n0, ptx::kEvictLast);
            if (++stage == C::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(2 * BM, BN, 1, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int rb = pair; rb < p.row_blocks; rb += num_pairs) {
        for (int nt = 0; nt < p.tiles_n; ++nt) {
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
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::mma_commit_pair(&empty[stage], 3);
            if (++stage == C::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          ptx::mma_commit_pair(&tfull[as], 3);
          if (++as == 2) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 0..7
    const int quad = warp & 3;
    const int grp = warp >> 2;
    uint8_t* stg = staging + warp * kStgBytes;
    const uint8_t* stg_partner = staging + (warp ^ 4) * kStgBytes;
    const int piece = lane & 7;   // 16-byte piece of a 128-byte chunk row handled by this lane
    const int rsub = lane >> 3;   // row within a group of 4 rows
    const uint32_t tempty_lead0 = ptx::mapa(&tempty[0], 0);
    const uint32_t tempty_lead1 = ptx::mapa(&tempty[1], 0);
    int as = 0;
    uint32_t aphase = 0;
    uint64_t pol_keep;  // the new residual is re-read by this same lane in pass 2: ask L2 to hold on to it
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));

    // x_old of the next chunk this warp will process, in the coalesced layout: rnext[it] = 4 columns of row it*4+rsub
    float4 rnext[8];
    float4 bnext = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prefetch = [&](int rb, int n0c) {
      const int mw = rb * (2 * BM) + static_cast<int>(rank) * BM + quad * 32;
      const int col = n0c + piece * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = mw + it * 4 + rsub;
        rnext[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.M) rnext[it] = *reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row) * p.ldr + col);
      }
      if (p.bias != nullptr) bnext = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    };
    if (pair < p.row_blocks && grp * CH < p.N) prefetch(pair, grp * CH);

    for (int rb = pair; rb < p.row_blocks; rb += num_pairs) {
      const int mw = rb * (2 * BM) + static_cast<int>(rank) * BM + quad * 32;  // first row of this warp
      float sk[8], ss[8], sq[8];  // shift, sum(x - shift), sum((x - shift)^2) per row slot
      int cnt = 0;                // values per row slot accumulated by this lane
      // ---------------------------------------------------------------- pass 1
      for (int nt = 0; nt < p.tiles_n; ++nt) {
        const int nt0 = nt * BN;
        ptx::mbar_wait(&tfull[as], aphase);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
#pragma unroll 1
        for (int c = grp; c < NCH; c += 2) {
          const int n0 = nt0 + c * CH;
          if (n0 >= p.N) break;
          float4 rcur[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) rcur[it] = rnext[it];
          const float4 bcur = bnext;
          {  // coordinates of the next chunk of this warp: same tile, next tile, or the next row block
            int c2 = c + 2, nt2 = nt, rb2 = rb;
            if (c2 >= NCH || nt0 + c2 * CH >= p.N) {
              c2 = grp;
              if (++nt2 == p.tiles_n) {
                nt2 = 0;
                rb2 += num_pairs;
              }
            }
            if (rb2 < p.row_blocks) prefetch(rb2, nt2 * BN + c2 * CH);
          }
          uint32_t r[32];
          ptx::tmem_ld_x32(t_row + c * CH, r);
          ptx::tmem_ld_wait();
          if (c + 2 >= NCH || n0 + 2 * CH >= p.N) {  // last chunk of this warp in the tile: the accumulator is drained
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(as == 0 ? tempty_lead0 : tempty_lead1);
          }
          {  // row-per-thread -> swizzled smem tile
            uint8_t* sb = stg + lane * 128;
            const int sw = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<uint4*>(sb + ((i ^ sw) << 4)) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          }
          __syncwarp();
          const int col = n0 + piece * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            float4 v = *reinterpret_cast<const float4*>(stg + rr * 128 + ((piece ^ (rr & 7)) << 4));
            v.x = (v.x + bcur.x) + rcur[it].x;  // same association as the unfused path: (acc + bias) + residual
            v.y = (v.y + bcur.y) + rcur[it].y;
            v.z = (v.z + bcur.z) + rcur[it].z;
            v.w = (v.w + bcur.w) + rcur[it].w;
            if (mw + rr < p.M) {
              float* dst = p.resid + static_cast<long long>(mw + rr) * p.ldr + col;
              asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                           "f"(v.w), "l"(pol_keep)
                           : "memory");
            }
            if (cnt == 0) {
              sk[it] = v.x;
              ss[it] = 0.f;
              sq[it] = 0.f;
            }
            const float d0 = v.x - sk[it], d1 = v.y - sk[it], d2 = v.z - sk[it], d3 = v.w - sk[it];
            ss[it] += (d0 + d1) + (d2 + d3);
            sq[it] = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq[it]))));
          }
          cnt += 4;
          __syncwarp();  // the tile is rewritten by the next chunk
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
      // ---------------------------------------------------------------- row statistics
      float mean[8], rstd[8];
      {
        const float nl = static_cast<float>(cnt);
        const float inv_nl = 1.0f / nl;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float n = nl;
          float m = fmaf(ss[it], inv_nl, sk[it]);
          float q = fmaf(-ss[it] * inv_nl, ss[it], sq[it]);
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {  // the 8 lanes of a row hold equal counts
            const float mb = __shfl_xor_sync(0xffffffffu, m, o);
            const float qb = __shfl_xor_sync(0xffffffffu, q, o);
            const float d = mb - m;
            m = fmaf(d, 0.5f, m);
            q = q + qb + d * d * n * 0.5f;
            n *= 2.0f;
          }
          mean[it] = m;
          rstd[it] = q;  // M2 for now
          if (piece == 0) {
            float* o = reinterpret_cast<float*>(stg) + (it * 4 + rsub) * 4;
            o[0] = n;
            o[1] = m;
            o[2] = q;
          }
        }
        named_bar_sync(1 + quad, 64);  // the two warps sharing these 32 rows
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float* o = reinterpret_cast<const float*>(stg_partner) + (it * 4 + rsub) * 4;
          float n = nl * 8.0f, m = mean[it], q = rstd[it];
          chan_combine(n, m, q, o[0], o[1], o[2]);
          mean[it] = m;
          rstd[it] = rsqrtf(fmaxf(q, 0.f) * p.inv_n + p.eps);
        }
        named_bar_sync(1 + quad, 64);  // partner has read my summary: the staging tile may be reused
      }
      // ---------------------------------------------------------------- pass 2: normalise what this lane stored
      for (int nt = 0; nt < p.tiles_n; ++nt) {
#pragma unroll 1
        for (int c = grp; c < NCH; c += 2) {
          const int n0 = nt * BN + c * CH;
          if (n0 >= p.N) break;
          const int col = n0 + piece * 4;
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + col));
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = mw + it * 4 + rsub;
            v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < p.M) v[it] = *reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row) * p.ldr + col);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = mw + it * 4 + rsub;
            const float m = mean[it], rs = rstd[it];
            const float o0 = fmaf((v[it].x - m) * rs, g.x, b.x);
            const float o1 = fmaf((v[it].y - m) * rs, g.y, b.y);
            const float o2 = fmaf((v[it].z - m) * rs, g.z, b.z);
            const float o3 = fmaf((v[it].w - m) * rs, g.w, b.w);
            if (row < p.M) {
              uint2 w;
              w.x = pack_bf16x2(o0, o1);
              w.y = pack_bf16x2(o2, o3);
              *reinterpret_cast<uint2*>(p.xn + static_cast<long long>(row) * p.ldxn + col) = w;
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<C::kTmemCols>(tmem_base);
  }
}

template <int BN>
int launch_ln(const CUtensorMap& tmA, const void* W, int64_t ldw, LnParams p, cudaStream_t stream) {
  using C = Cfg3<BN>;
  auto kern = gemm_pair_ln_kernel<BN>;
  static int configured_dev = -1;
  static int max_pairs = 0;
  int dev = 0;
  EVT_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    EVT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() & ~1, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    int n = 0;
    EVT_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    if (n <= 0) return fail(EVT_ERR_CUDA, "gemm+ln: no CTA pair of this configuration fits on the device");
    max_pairs = n;
    configured_dev = dev;
  }
  CUtensorMap tmW;
  int rc = make_tmap_2d(&tmW, W, 2, static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.K), static_cast<uint64_t>(ldw), BN / 2, 64);
  if (rc != EVT_OK) return rc;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int pairs = p.row_blocks < max_pairs ? p.row_blocks : max_pairs;
  EVT_CUDA(launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), C::kSmemBytes, stream, pdl_for_rows(p.M), tmA, tmW, p));
  EVT_LAUNCH_CHECK("gemm_pair_ln_kernel");
  return EVT_OK;
}

}  // namespace

bool gemm_res_ln_supported(int64_t M, int N, int K) {
  // full 64-column pairs of chunks per row (both warps of a quadrant get the same share), rows fit the prefetch math
  return N >= 64 && N <= 1024 && N % 64 == 0 && K > 0 && M > 0;
}

int gemm_res_ln_launch(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float* resid, int64_t ldr,
                       const float* gamma, const float* beta, float eps, void* xn, int64_t ldxn, int64_t M, int N, int K,
                       cudaStream_t stream) {
  EVT_CHECK_ARG(A && W && resid && gamma && beta && xn, "gemm+ln: null pointer");
  if (!gemm_res_ln_supported(M, N, K)) return fail(EVT_ERR_UNSUPPORTED, "gemm+ln: N must be a multiple of 64 in [64, 1024]");
  EVT_CHECK_ARG(M < (1ll << 31) - 512, "gemm+ln: M too large");
  EVT_CHECK_ARG(lda >= K && ldw >= K && ldr >= N && ldxn >= N, "gemm+ln: leading dimension smaller than the row length");
  EVT_CHECK_ARG(reinterpret_cast<uintptr_t>(resid) % 16 == 0 && ldr % 4 == 0, "gemm+ln: residual rows must be 16-byte aligned");
  EVT_CHECK_ARG(reinterpret_cast<uintptr_t>(xn) % 8 == 0 && ldxn % 4 == 0, "gemm+ln: xn rows must be 8-byte aligned");
  EVT_CHECK_ARG((bias == nullptr || reinterpret_cast<uintptr_t>(bias) % 16 == 0) && reinterpret_cast<uintptr_t>(gamma) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(beta) % 16 == 0,
                "gemm+ln: bias / gamma / beta must be 16-byte aligned");
  CUtensorMap tmA;
  int rc = make_tmap_2d(&tmA, A, 2, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), BM, 64);
  if (rc != EVT_OK) return rc;
  LnParams p;
  p.bias = bias;
  p.resid = resid;
  p.gamma = gamma;
  p.beta = beta;
  p.xn = reinterpret_cast<__nv_bfloat16*>(xn);
  p.ldr = ldr;
  p.ldxn = ldxn;
  p.M = static_cast<int>(M);
  p.N = N;
  p.K = K;
  p.num_kb = (K + 63) / 64;
  p.row_blocks = static_cast<int>((M + 2 * BM - 1) / (2 * BM));
  p.eps = eps;
  p.inv_n = 1.0f / static_cast<float>(N);
  p.tiles_n = 0;
  // tile width: the widest of 256 / 192 / 128 that wastes no columns (768 -> 3 x 256, 384 -> 2 x 192, 192 -> 192)
  int bn = 256;
  long best = -1;
  for (int cand : {256, 192, 128}) {
    const long cost = static_cast<long>((N + cand - 1) / cand) * cand;
    if (best < 0 || cost < best) {
      best = cost;
      bn = cand;
    }
  }
  switch (bn) {
    case 256: return launch_ln<256>(tmA, W, ldw, p, stream);
    case 192: return launch_ln<192>(tmA, W, ldw, p, stream);
    default: return launch_ln<128>(tmA, W, ldw, p, stream);
  }
}

}  // namespace evt

/* x <- x + A W^T + bias (f32, in place); xn <- LayerNorm(x) gamma + beta (bf16).  See include/evt.h. */
extern "C" int evt_gemm_residual_layernorm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                           float* resid, int64_t ldr, const float* gamma, const float* beta, float eps,
                                           void* xn, int64_t ldxn, int64_t M, int N, int K, evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  return evt::gemm_res_ln_launch(A, lda, W, ldw, bias, resid, ldr, gamma, beta, eps, xn, ldxn, M, N, K,
                                 static_cast<cudaStream_t>(stream));
}
