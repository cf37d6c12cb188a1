// Global-pool attention pieces of the ContextPath / FeatureFusionModule and the nearest
// up-sampling between context stages.
//   reference: AttentionRefinementModule.forward (model/model_stages.py:77-85),
//   ContextPath.forward (model_stages.py:120-133), FeatureFusionModule.forward (175-185).
// Full-map passes are HBM-bound and work on 8-channel (16 B) vectors; the per-image [N, C]
// vectors go through the tiny dense layers below in fp32.
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

__device__ __forceinline__ void ld8a(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8a(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                            pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
// PyTorch 'nearest': src = min(floor(dst * (in/out)), in-1), evaluated in fp32.
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  const int s = (int)floorf(dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

// out[n][c] += sum over the pixels of image n of x[n, pixel, c]      (grid: x = chunk, y = n)
__global__ void __launch_bounds__(256)
pool_sum_kernel(const __nv_bfloat16* __restrict__ x, int ld, int HW, int C, float* __restrict__ out) {
  extern __shared__ float s_acc[];  // [C]
  const int groups = C >> 3;
  const int py = blockDim.x / groups;
  const int g = threadIdx.x % groups, ty = threadIdx.x / groups;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float acc[8] = {0};
  const __nv_bfloat16* base = x + (int64_t)n * HW * ld;
  for (int p = blockIdx.x * py + ty; p < HW; p += gridDim.x * py) {
    float v[8];
    ld8a(base + (int64_t)p * ld + g * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[g * 8 + j], acc[j]);
  __syncthreads();
  flush_add_v4(out + (int64_t)n * C, s_acc, C, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------ tiny dense layers
constexpr int kMaxBatch = 64;

// These layers move a few KB and are pure latency chains, so each kernel is arranged to keep many
// independent loads in flight (the first versions walked batch x channel serially and took 10-60 us
// per launch under ncu): a lane carries the accumulators of kFcNB batch rows at once, loops are
// unrolled over independent channels, per-row statistics come from shared memory.
constexpr int kFcNB = 8;

// pre[n][co] = in_scale * sum_ci in[n][ci] * W[co][ci];  optional BatchNorm over the batch
// (train: batch statistics + running update, eval: running statistics); act 0 none, 1 relu,
// 3 sigmoid.  One CTA per output channel (4 warps split the reduction over the input channels).
__global__ void __launch_bounds__(128)
fc_small_fwd_kernel(const float* __restrict__ in, float in_scale, int N, int Cin, int Co,
                    const float* __restrict__ W, int has_bn, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ running_mean,
                    float* __restrict__ running_var, float momentum, float eps, int training,
                    int act, float* __restrict__ pre, float* __restrict__ out,
                    float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ float s_part[4][kMaxBatch];
  __shared__ float s_y[1][kMaxBatch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x;       // one CTA per output channel; its 4 warps split the input channels
  const float* wr = W + (int64_t)co * Cin;
  for (int n0 = 0; n0 < N; n0 += kFcNB) {
    float acc[kFcNB];
#pragma unroll
    for (int j = 0; j < kFcNB; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (int ci = threadIdx.x; ci < Cin; ci += 128) {
      const float w = __ldg(wr + ci);
#pragma unroll
      for (int j = 0; j < kFcNB; ++j)
        if (n0 + j < N) acc[j] += __ldg(in + (int64_t)(n0 + j) * Cin + ci) * w;
    }
#pragma unroll
    for (int j = 0; j < kFcNB; ++j) {
      const float t = warp_sum(acc[j]);
      if (lane == 0 && n0 + j < N) s_part[warp][n0 + j] = t;
    }
  }
  __syncthreads();
  if (warp != 0) return;
  for (int n = lane; n < N; n += 32)
    s_y[0][n] = (s_part[0][n] + s_part[1][n] + s_part[2][n] + s_part[3][n]) * in_scale;
  __syncwarp();
  float mean = 0.f, rstd = 1.f, g = 1.f, b = 0.f;
  if (has_bn) {
    g = gamma[co];
    b = beta[co];
    if (training) {
      float s1 = 0.f, s2 = 0.f;
      for (int n = 0; n < N; ++n) s1 += s_y[0][n];
      mean = s1 / N;
      for (int n = 0; n < N; ++n) {
        const float d = s_y[0][n] - mean;
        s2 += d * d;
      }
      const float var = s2 / N;
      rstd = rsqrtf(var + eps);
      if (lane == 0 && running_mean != nullptr) {
        running_mean[co] = (1.f - momentum) * running_mean[co] + momentum * mean;
        running_var[co] = (1.f - momentum) * running_var[co] + momentum * (N > 1 ? var * N / (N - 1) : var);
      }
    } else {
      mean = running_mean[co];
      rstd = rsqrtf(running_var[co] + eps);
    }
    if (lane == 0 && mean_out != nullptr) {
      mean_out[co] = mean;
      rstd_out[co] = rstd;
    }
  }
  for (int n = lane; n < N; n += 32) {
    const float y = s_y[0][n];
    float t = has_bn ? (y - mean) * rstd * g + b : y;
    if (act == 1) t = fmaxf(t, 0.f);
    else if (act == 3) t = 1.f / (1.f + __expf(-t));
    if (pre != nullptr) pre[(int64_t)n * Co + co] = y;
    out[(int64_t)n * Co + co] = t;
  }
}

// Backward, part A (one CTA per output channel): activation + BatchNorm backward over the batch,
// dgamma / dbeta, dpre[n][co], and dW[co][ci] += in_scale * sum_n dpre[n][co] * in[n][ci].  Every warp
// derives the (tiny) per-row gradients itself; the four warps then split the input channels of dW.
__global__ void __launch_bounds__(128)
fc_small_bwd_a_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                      const float* __restrict__ pre, const float* __restrict__ in, float in_scale,
                      int N, int Cin, int Co, int has_bn, int training,
                      const float* __restrict__ gamma, const float* __restrict__ mean,
                      const float* __restrict__ rstd, int act, float* __restrict__ dpre,
                      float* __restrict__ dW, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s_g[4][kMaxBatch], s_x[4][kMaxBatch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x;
  const float mu = has_bn ? mean[co] : 0.f, rs = has_bn ? rstd[co] : 1.f, gm = has_bn ? gamma[co] : 1.f;
  float sg = 0.f, sgx = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float o = out[(int64_t)n * Co + co];
    float gg = dout[(int64_t)n * Co + co];
    if (act == 1) gg = o > 0.f ? gg : 0.f;
    else if (act == 3) gg *= o * (1.f - o);
    const float xh = has_bn ? (pre[(int64_t)n * Co + co] - mu) * rs : 0.f;
    s_g[warp][n] = gg;
    s_x[warp][n] = xh;
    sg += gg;
    sgx += gg * xh;
  }
  if (has_bn) {
    sg = warp_sum(sg);
    sgx = warp_sum(sgx);
    for (int n = lane; n < N; n += 32) {
      const float gg = s_g[warp][n];
      s_g[warp][n] = training ? gm * rs * (gg - sg / N - s_x[warp][n] * sgx / N) : gm * rs * gg;
    }
    if (warp == 0 && lane == 0) {
      if (dgamma != nullptr) dgamma[co] += sgx;
      if (dbeta != nullptr) dbeta[co] += sg;
    }
  }
  __syncwarp();
  if (warp == 0)
    for (int n = lane; n < N; n += 32) dpre[(int64_t)n * Co + co] = s_g[0][n];
  if (dW != nullptr) {
    float* dwr = dW + (int64_t)co * Cin;
#pragma unroll 4
    for (int ci = threadIdx.x; ci < Cin; ci += 128) {
      const float old = dwr[ci];
      float acc = 0.f;
      for (int n0 = 0; n0 < N; n0 += kFcNB) {
#pragma unroll
        for (int j = 0; j < kFcNB; ++j)
          if (n0 + j < N) acc += s_g[warp][n0 + j] * __ldg(in + (int64_t)(n0 + j) * Cin + ci);
      }
      dwr[ci] = old + acc * in_scale;
    }
  }
}

// Backward, part B: din[n][ci] (+)= in_scale * sum_co dpre[n][co] * W[co][ci].
// Block = 16 input channels x 16 slices of the output channels (short dependent chains even for
// Co = 256); a thread carries kFcNB batch rows; the slices are combined through shared memory.
__global__ void __launch_bounds__(256)
fc_small_bwd_b_kernel(const float* __restrict__ dpre, const float* __restrict__ W, float in_scale,
                      int N, int Cin, int Co, float* __restrict__ din, int accumulate) {
  __shared__ float s_red[16][kFcNB][17];
  const int cl = threadIdx.x & 15, cg = threadIdx.x >> 4;
  const int ci = blockIdx.x * 16 + cl;
  for (int n0 = 0; n0 < N; n0 += kFcNB) {
    float acc[kFcNB];
#pragma unroll
    for (int j = 0; j < kFcNB; ++j) acc[j] = 0.f;
    if (ci < Cin) {
#pragma unroll 4
      for (int co = cg; co < Co; co += 16) {
        const float w = __ldg(W + (int64_t)co * Cin + ci);
#pragma unroll
        for (int j = 0; j < kFcNB; ++j)
          if (n0 + j < N) acc[j] += __ldg(dpre + (int64_t)(n0 + j) * Co + co) * w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kFcNB; ++j) s_red[cg][j][cl] = acc[j];
    __syncthreads();
    if (threadIdx.x < kFcNB * 16) {
      const int j = threadIdx.x >> 4, c = threadIdx.x & 15;
      const int n = n0 + j, cc = blockIdx.x * 16 + c;
      if (n < N && cc < Cin) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) v += s_red[k][j][c];
        v *= in_scale;
        float* d = din + (int64_t)n * Cin + cc;
        *d = accumulate ? *d + v : v;
      }
    }
  }
}

// ---------------------------------------------------------- broadcast scale/add
// out[n, ho, wo, c] = a[n, hs, ws, c] * (s[n, c] + s_plus) + (v ? v[n, c] * v_scale : 0) + (t ? t[n, hs, ws, c] : 0)
// with (hs, ws) the nearest-neighbour source of (ho, wo) (identity when the sizes agree).
__global__ void __launch_bounds__(256)
scale_add_bcast_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, int Hs, int Ws,
                       const float* __restrict__ s, float s_plus, const float* __restrict__ v,
                       float v_scale, const __nv_bfloat16* __restrict__ t, int t_ld,
                       __nv_bfloat16* __restrict__ out, int out_ld, int N, int Ho, int Wo, int C) {
  const int groups = C >> 3;
  const uint32_t total = (uint32_t)N * Ho * Wo * groups;   // host-checked < 2^31: 32-bit index math
  const float sh = (float)Hs / (float)Ho, sw = (float)Ws / (float)Wo;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const uint32_t p = i / groups;
    const int wo = (int)(p % Wo);
    const uint32_t t2 = p / Wo;
    const int ho = (int)(t2 % Ho);
    const int n = (int)(t2 / Ho);
    const int hs = (Hs == Ho) ? ho : nearest_src(ho, sh, Hs);
    const int ws = (Ws == Wo) ? wo : nearest_src(wo, sw, Ws);
    const int64_t q = ((int64_t)n * Hs + hs) * Ws + ws;
    float x[8], o[8];
    ld8a(a + q * a_ld + g * 8, x);
    float sv[8], vv[8];
    if (s != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + (int64_t)n * C + g * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(s + (int64_t)n * C + g * 8) + 1);
      sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
    }
    if (v != nullptr) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(v + (int64_t)n * C + g * 8));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(v + (int64_t)n * C + g * 8) + 1);
      vv[0] = v0.x; vv[1] = v0.y; vv[2] = v0.z; vv[3] = v0.w; vv[4] = v1.x; vv[5] = v1.y; vv[6] = v1.z; vv[7] = v1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r = x[j] * ((s != nullptr ? sv[j] : 0.f) + s_plus);
      if (v != nullptr) r += vv[j] * v_scale;
      o[j] = r;
    }
    if (t != nullptr) {
      float y[8];
      ld8a(t + q * t_ld + g * 8, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += y[j];
    }
    st8a(out + (int64_t)p * out_ld + g * 8, o);
  }
}

// Backward of the above w.r.t. the source-resolution sum:
//   dsum[n, hs, ws, c] = sum over the nearest-neighbour footprint of dout[n, ho, wo, c]
//   (stored when dsum != nullptr), dot[n][c] += sum_{hs,ws} dsum * b[n, hs, ws, c]
//   (when b != nullptr) and vsum[n][c] += sum_{hs,ws} dsum (when vsum != nullptr).
// grid: x = chunk of source pixels, y = n.
__global__ void __launch_bounds__(256)
upsum_dot_reduce_kernel(const __nv_bfloat16* __restrict__ dout, int dout_ld, int Ho, int Wo,
                        const __nv_bfloat16* __restrict__ b, int b_ld,
                        __nv_bfloat16* __restrict__ dsum, int dsum_ld, int Hs, int Ws, int C,
                        float* __restrict__ dot, float* __restrict__ vsum) {
  extern __shared__ float s_acc[];  // [2][C]
  const int groups = C >> 3;
  const int py = blockDim.x / groups;
  const int g = threadIdx.x % groups, ty = threadIdx.x / groups;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const float sh = (float)Hs / (float)Ho, sw = (float)Ws / (float)Wo;
  float a1[8] = {0}, a2[8] = {0};
  for (int p = blockIdx.x * py + ty; p < Hs * Ws; p += gridDim.x * py) {
    const int hs = p / Ws, ws = p - hs * Ws;
    float acc[8] = {0};
    if (Hs == Ho && Ws == Wo) {
      ld8a(dout + ((int64_t)n * Ho * Wo + p) * dout_ld + g * 8, acc);
    } else {
      int h_lo = (int)floorf(hs / sh) - 1;
      if (h_lo < 0) h_lo = 0;
      int w_lo = (int)floorf(ws / sw) - 1;
      if (w_lo < 0) w_lo = 0;
      for (int ho = h_lo; ho < Ho; ++ho) {
        const int sh_i = nearest_src(ho, sh, Hs);
        if (sh_i < hs) continue;
        if (sh_i > hs) break;
        for (int wo = w_lo; wo < Wo; ++wo) {
          const int sw_i = nearest_src(wo, sw, Ws);
          if (sw_i < ws) continue;
          if (sw_i > ws) break;
          float v[8];
          ld8a(dout + (((int64_t)n * Ho + ho) * Wo + wo) * dout_ld + g * 8, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
      }
    }
    const int64_t q = (int64_t)n * Hs * Ws + p;
    if (dsum != nullptr) st8a(dsum + q * dsum_ld + g * 8, acc);
    if (b != nullptr) {
      float v[8];
      ld8a(b + q * b_ld + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) a1[j] += acc[j] * v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) a2[j] += acc[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_acc[g * 8 + j], a1[j]);
    atomicAdd(&s_acc[C + g * 8 + j], a2[j]);
  }
  __syncthreads();
  if (dot != nullptr) flush_add_v4(dot + (int64_t)n * C, s_acc, C, threadIdx.x, blockDim.x);
  if (vsum != nullptr) flush_add_v4(vsum + (int64_t)n * C, s_acc + C, C, threadIdx.x, blockDim.x);
}

static int thr_for(int C) {
  const int groups = C / 8;
  int t = (256 / groups) * groups;
  return t == 0 ? groups : t;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_pool_sum(const void* x, int ld, int N, int HW, int C, float* out, cudaStream_t stream) {
  if (C % 8 || C > 2048) return set_error(B200_EINVAL, "pool_sum: C=%d must be a multiple of 8 (<= 2048)", C);
  const int threads = thr_for(C);
  const int py = threads / (C / 8);
  int chunks = (HW + py * 8 - 1) / (py * 8);
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  pool_sum_kernel<<<dim3(chunks, N), threads, C * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(x), ld, HW, C, out);
  return check_launch("pool_sum");
}

int b200_fc_small_fwd(const float* in, float in_scale, int N, int Cin, int Co, const float* W,
                      int has_bn, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, int training, int act,
                      float* pre, float* out, float* mean_out, float* rstd_out, cudaStream_t stream) {
  if (N > kMaxBatch) return set_error(B200_EINVAL, "fc_small: batch %d > %d", N, kMaxBatch);
  fc_small_fwd_kernel<<<Co, 128, 0, stream>>>(in, in_scale, N, Cin, Co, W, has_bn, gamma, beta,
                                                        running_mean, running_var, momentum, eps,
                                                        training, act, pre, out, mean_out, rstd_out);
  return check_launch("fc_small_fwd");
}

int b200_fc_small_bwd(const float* dout, const float* out, const float* pre, const float* in,
                      float in_scale, int N, int Cin, int Co, const float* W, int has_bn,
                      int training, const float* gamma, const float* mean, const float* rstd,
                      int act, float* dpre_scratch, float* dW, float* dgamma, float* dbeta,
                      float* din, int accumulate_din, cudaStream_t stream) {
  if (N > kMaxBatch) return set_error(B200_EINVAL, "fc_small: batch %d > %d", N, kMaxBatch);
  fc_small_bwd_a_kernel<<<Co, 128, 0, stream>>>(dout, out, pre, in, in_scale, N, Cin, Co, has_bn,
                                                          training, gamma, mean, rstd, act,
                                                          dpre_scratch, dW, dgamma, dbeta);
  int rc = check_launch("fc_small_bwd_a");
  if (rc) return rc;
  if (din != nullptr) {
    fc_small_bwd_b_kernel<<<(Cin + 15) / 16, 256, 0, stream>>>(dpre_scratch, W, in_scale, N, Cin, Co, din, accumulate_din);
    rc = check_launch("fc_small_bwd_b");
  }
  return rc;
}

int b200_scale_add_bcast(const void* a, int a_ld, int Hs, int Ws, const float* s, float s_plus,
                         const float* v, float v_scale, const void* t, int t_ld, void* out,
                         int out_ld, int N, int Ho, int Wo, int C, cudaStream_t stream) {
  if (C % 8) return set_error(B200_EINVAL, "scale_add_bcast: C=%d must be a multiple of 8", C);
  int64_t total = (int64_t)N * Ho * Wo * (C / 8);
  if (total >= (1ll << 31)) return set_error(B200_EINVAL, "scale_add_bcast: tensor too large");
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  scale_add_bcast_kernel<<<(int)blocks, 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(a), a_ld, Hs, Ws, s, s_plus, v, v_scale,
      static_cast<const __nv_bfloat16*>(t), t_ld, static_cast<__nv_bfloat16*>(out), out_ld, N, Ho, Wo, C);
  return check_launch("scale_add_bcast");
}

int b200_upsum_dot_reduce(const void* dout, int dout_ld, int Ho, int Wo, const void* b, int b_ld,
                          void* dsum, int dsum_ld, int N, int Hs, int Ws, int C, float* dot,
                          float* vsum, cudaStream_t stream) {
  if (C % 8 || C > 2048) return set_error(B200_EINVAL, "upsum_dot_reduce: C=%d must be a multiple of 8 (<= 2048)", C);
  const int threads = thr_for(C);
  const int py = threads / (C / 8);
  int chunks = (Hs * Ws + py * 4 - 1) / (py * 4);
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  upsum_dot_reduce_kernel<<<dim3(chunks, N), threads, 2 * C * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(dout), dout_ld, Ho, Wo, static_cast<const __nv_bfloat16*>(b), b_ld,
      static_cast<__nv_bfloat16*>(dsum), dsum_ld, Hs, Ws, C, dot, vsum);
  return check_launch("upsum_dot_reduce");
}

}  // extern "C"
