// T2T-ViT front-end kernels (modeling/models/t2t_vit.py:43-88, modeling/layers/transformer_encoder.py:39-101).
//
//   unfold_ln     soft split (tf_Unfold, channel-last, depth order (kh,kw,c), zero pad) fused with the performer's
//                 first LayerNorm: one warp gathers one window, normalises it in registers and writes the bf16 row
//                 that feeds the kqv GEMM -- the unfolded tensor never exists un-normalised in HBM.
//   performer     linear attention with positive random features (m = 32, emb = 64):
//                   kp = exp(w k - |k|^2/2)/sqrt(m);  ksum = sum_t kp;  kptv = sum_t v kp^T      (reduce, per image)
//                   qp = exp(w q - |q|^2/2)/sqrt(m);  y = (qp kptv^T) / (qp.ksum + 1e-8)          (apply, per token)
//                 The reduction over tokens is two-level with a fixed summation order (deterministic, no atomics).
// The dense layers around them (kqv, attn_output, MLP, project) run on the tcgen05 GEMM with fused epilogues.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace evt {
namespace {

constexpr int kEmb = 64;
constexpr int kM = 32;
constexpr int kChunk = 196;  // tokens per reduce block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- unfold (+ LN)
// One warp per output row.  Row length L = k*k*C <= 32 * kMaxPerLane.
constexpr int kMaxPerLane = 18;  // 576 / 32

template <typename TIN, bool LN>
__global__ void __launch_bounds__(256) unfold_ln_kernel(const TIN* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        long long ldo, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int B, int H, int W,
                                                        int C, int k, int s, int p, int oh, int ow, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int L = k * k * C;
  const int ox = static_cast<int>(row % ow);
  const int oy = static_cast<int>((row / ow) % oh);
  const long long b = row / (static_cast<long long>(ow) * oh);
  const int kC = k * C;
  float v[kMaxPerLane];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int col = i * 32 + lane;
    float val = 0.f;
    if (col < L) {
      const int ky = col / kC;
      const int rem = col - ky * kC;  // kx*C + c : contiguous in the input row
      const int kx = rem / C;
      const int iy = oy * s - p + ky, ix = ox * s - p + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        val = static_cast<float>(x[((b * H + iy) * W + ox * s - p) * C + rem]);
    }
    v[i] = val;
    sum += val;
  }
  float mean = 0.f, rstd = 1.f;
  if (LN) {
    mean = warp_sum(sum) / static_cast<float>(L);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int col = i * 32 + lane;
      if (col < L) {
        const float d = v[i] - mean;
        q += d * d;
      }
    }
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(L) + eps);
  }
  __nv_bfloat16* orow = out + row * ldo;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int col = i * 32 + lane;
    if (col < ldo) {
      float o = 0.f;
      if (col < L) o = LN ? (v[i] - mean) * rstd * gamma[col] + beta[col] : v[i];
      orow[col] = __float2bfloat16_rn(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------- performer
// prm_exp for one token held by a warp: lane m returns exp(w[m,:].x - |x|^2/2) / sqrt(32).
// xa, xb = elements 2*lane, 2*lane+1 of the 64-vector; wreg = row `lane` of w.
__device__ __forceinline__ float prm_exp_lane(const float (&wreg)[kEmb], float xa, float xb) {
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float a = __shfl_sync(0xffffffffu, xa, i);
    const float b = __shfl_sync(0xffffffffu, xb, i);
    dot = fmaf(wreg[2 * i], a, dot);
    dot = fmaf(wreg[2 * i + 1], b, dot);
  }
  const float xd = 0.5f * warp_sum(xa * xa + xb * xb);
  return __expf(dot - xd) * 0.17677669529663687f;  // 1/sqrt(32)
}

// kqv: bf16 [B*T, ld] with k at cols [0,64), q at [64,128), v at [128,192).
// partial: f32 [B, nsplit, 32 + 64*32]  (ksum | kptv[n][m])
__global__ void __launch_bounds__(128) performer_reduce_kernel(const __nv_bfloat16* __restrict__ kqv, long long ld,
                                                               const float* __restrict__ w, float* __restrict__ partial,
                                                               int T, int nsplit) {
  __shared__ float red[4][kM + kEmb * kM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, b = blockIdx.y;
  float wreg[kEmb];
#pragma unroll
  for (int i = 0; i < kEmb; ++i) wreg[i] = w[lane * kEmb + i];
  float ksum = 0.f;
  float acc[kEmb];  // kptv[n][lane]
#pragma unroll
  for (int n = 0; n < kEmb; ++n) acc[n] = 0.f;
  const int t0 = split * kChunk, t1 = min(T, t0 + kChunk);
  for (int t = t0 + warp; t < t1; t += 4) {
    const __nv_bfloat16* rowp = kqv + (static_cast<long long>(b) * T + t) * ld;
    const float2 kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rowp + 2 * lane));
    const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rowp + 128 + 2 * lane));
    const float kp = prm_exp_lane(wreg, kk.x, kk.y);
    ksum += kp;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float va = __shfl_sync(0xffffffffu, vv.x, i);
      const float vb = __shfl_sync(0xffffffffu, vv.y, i);
      acc[2 * i] = fmaf(va, kp, acc[2 * i]);
      acc[2 * i + 1] = fmaf(vb, kp, acc[2 * i + 1]);
    }
  }
  red[warp][lane] = ksum;
#pragma unroll
  for (int n = 0; n < kEmb; ++n) red[warp][kM + n * kM + lane] = acc[n];
  __syncthreads();
  float* dst = partial + (static_cast<long long>(b) * nsplit + split) * (kM + kEmb * kM);
  for (int i = threadIdx.x; i < kM + kEmb * kM; i += 128) dst[i] = ((red[0][i] + red[1][i]) + red[2][i]) + red[3][i];
}

__global__ void __launch_bounds__(256) performer_reduce_final_kernel(const float* __restrict__ partial,
                                                                     float* __restrict__ stats, int nsplit) {
  const int b = blockIdx.x;
  const int n = kM + kEmb * kM;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[(static_cast<long long>(b) * nsplit + k) * n + i];
    stats[static_cast<long long>(b) * n + i] = s;
  }
}

// per token: yattn (bf16 [B*T, 64]) = (qp kptv^T)/(qp.ksum + eps);  vout (f32 [B*T, 64]) = v (the skip connection that
// the attn_output GEMM then reduce-adds into, transformer_encoder.py:93).
__global__ void __launch_bounds__(128) performer_apply_kernel(const __nv_bfloat16* __restrict__ kqv, long long ld,
                                                              const float* __restrict__ w, const float* __restrict__ stats,
                                                              __nv_bfloat16* __restrict__ yattn, float* __restrict__ vout,
                                                              int T, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const float* st = stats + static_cast<long long>(b) * (kM + kEmb * kM);
  float wreg[kEmb];
#pragma unroll
  for (int i = 0; i < kEmb; ++i) wreg[i] = w[lane * kEmb + i];
  const float ksum = st[lane];
  float kv0[kM], kv1[kM];  // kptv[lane][m], kptv[lane+32][m]
#pragma unroll
  for (int m = 0; m < kM; ++m) {
    kv0[m] = st[kM + lane * kM + m];
    kv1[m] = st[kM + (lane + 32) * kM + m];
  }
  const int t0 = blockIdx.x * kChunk, t1 = min(T, t0 + kChunk);
  for (int t = t0 + warp; t < t1; t += 4) {
    const long long r = static_cast<long long>(b) * T + t;
    const __nv_bfloat16* rowp = kqv + r * ld;
    const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rowp + 64 + 2 * lane));
    const float qp = prm_exp_lane(wreg, qq.x, qq.y);
    const float D = warp_sum(qp * ksum);
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int m = 0; m < kM; ++m) {
      const float qm = __shfl_sync(0xffffffffu, qp, m);
      y0 = fmaf(qm, kv0[m], y0);
      y1 = fmaf(qm, kv1[m], y1);
    }
    const float inv = 1.0f / (D + eps);
    yattn[r * kEmb + lane] = __float2bfloat16_rn(y0 * inv);
    yattn[r * kEmb + lane + 32] = __float2bfloat16_rn(y1 * inv);
    vout[r * kEmb + lane] = __bfloat162float(rowp[128 + lane]);
    vout[r * kEmb + lane + 32] = __bfloat162float(rowp[128 + lane + 32]);
  }
}

}  // namespace

int unfold_ln_launch(const void* x, int x_dtype, void* out, int64_t ldo, const float* gamma, const float* beta, float eps,
                     int B, int H, int W, int C, int k, int s, int p, cudaStream_t st) {
  EVT_CHECK_ARG(x && out, "unfold_ln: null pointer");
  EVT_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && k > 0 && s > 0 && p >= 0, "unfold_ln: bad sizes");
  const int L = k * k * C;
  EVT_CHECK_ARG(ldo >= L, "unfold_ln: ldo smaller than k*k*C");
  if (ldo > 32 * kMaxPerLane) return fail(EVT_ERR_UNSUPPORTED, "unfold_ln: rows longer than 576 elements are not implemented");
  EVT_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "unfold_ln: gamma and beta must both be given or both be null");
  const int oh = (H + 2 * p - k) / s + 1, ow = (W + 2 * p - k) / s + 1;
  EVT_CHECK_ARG(oh > 0 && ow > 0, "unfold_ln: empty output");
  const long long rows = static_cast<long long>(B) * oh * ow;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const bool ln = gamma != nullptr;
  if (x_dtype == EVT_F32) {
    const float* xi = reinterpret_cast<const float*>(x);
    if (ln) unfold_ln_kernel<float, true><<<grid, 256, 0, st>>>(xi, o, ldo, gamma, beta, eps, B, H, W, C, k, s, p, oh, ow, rows);
    else unfold_ln_kernel<float, false><<<grid, 256, 0, st>>>(xi, o, ldo, gamma, beta, eps, B, H, W, C, k, s, p, oh, ow, rows);
  } else if (x_dtype == EVT_BF16) {
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
    if (ln) unfold_ln_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(xi, o, ldo, gamma, beta, eps, B, H, W, C, k, s, p, oh, ow, rows);
    else unfold_ln_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(xi, o, ldo, gamma, beta, eps, B, H, W, C, k, s, p, oh, ow, rows);
  } else {
    return fail(EVT_ERR_INVALID, "unfold_ln: x dtype must be f32 or bf16");
  }
  EVT_LAUNCH_CHECK("unfold_ln");
  return EVT_OK;
}

}  // namespace evt

using namespace evt;

extern "C" int evt_unfold_ln_nhwc(const void* x, int x_dtype, void* out, int64_t ldo, const float* gamma,
                                  const float* beta, float eps, int B, int H, int W, int C, int k, int s, int p,
                                  evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  return unfold_ln_launch(x, x_dtype, out, ldo, gamma, beta, eps, B, H, W, C, k, s, p, static_cast<cudaStream_t>(stream));
}

extern "C" int evt_performer_workspace_bytes(int B, int T, size_t* out) {
  EVT_CHECK_ARG(out != nullptr && B > 0 && T > 0, "performer_workspace_bytes: bad arguments");
  const size_t nsplit = (T + kChunk - 1) / kChunk;
  *out = static_cast<size_t>(B) * (nsplit + 1) * (kM + kEmb * kM) * sizeof(float);
  return EVT_OK;
}

extern "C" int evt_performer_fwd(const void* kqv, int64_t ld, const float* w, void* yattn, float* vout, void* workspace,
                                 int B, int T, int emb, int m, float eps, evt_stream stream) {
  int rc = evt_device_check();
  if (rc != EVT_OK) return rc;
  EVT_CHECK_ARG(kqv && w && yattn && vout && workspace, "performer: null pointer");
  EVT_CHECK_ARG(B > 0 && T > 0 && B <= 65535, "performer: B in 1..65535 and T > 0");
  if (emb != kEmb || m != kM) return fail(EVT_ERR_UNSUPPORTED, "performer: only emb = 64, m = 32 (T2T token_size 64, kernel_ratio 0.5) is implemented");
  EVT_CHECK_ARG(ld >= 3 * kEmb && ld % 2 == 0, "performer: kqv leading dimension must be >= 192 and even");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nsplit = (T + kChunk - 1) / kChunk;
  float* partial = reinterpret_cast<float*>(workspace);
  float* stats = partial + static_cast<size_t>(B) * nsplit * (kM + kEmb * kM);
  dim3 grid(nsplit, B);
  performer_reduce_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(kqv), ld, w, partial, T, nsplit);
  EVT_LAUNCH_CHECK("performer_reduce");
  performer_reduce_final_kernel<<<B, 256, 0, st>>>(partial, stats, nsplit);
  EVT_LAUNCH_CHECK("performer_reduce_final");
  performer_apply_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(kqv), ld, w, stats,
                                               reinterpret_cast<__nv_bfloat16*>(yattn), vout, T, eps);
  EVT_LAUNCH_CHECK("performer_apply");
  return EVT_OK;
}
