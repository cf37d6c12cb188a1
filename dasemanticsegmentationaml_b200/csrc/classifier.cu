// The discriminators' last layer: Conv2d(512, 1, kernel 4, stride 2, pad 1) with bias — an
// N = 1 "GEMV" convolution (0.07 GFLOP at batch 8), run as one warp per output pixel.
//   reference: self.classifier in model/discriminator.py:13,47,97 (forward at 26,72,132)
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

// All three kernels stage the filter transposed in shared memory: s_w[tap][c] with a pitch of
// C + 4 floats (keeps float4 alignment, spreads the transposing writes over banks).  They are
// latency-bound GEMV-like passes over a 17 MB tensor, so every kernel issues all loads of a work
// item before the first dependent FMA and is launched as ~2 CTAs per SM that loop over their items
// (the first versions took 30 / 66 / 82 us inside the step's graph).
__device__ __forceinline__ void stage_filter_t(const float* __restrict__ w, int C, float* s_w) {
  const int pitch = C + 4;
#pragma unroll 8
  for (int i = threadIdx.x; i < 16 * C; i += blockDim.x) s_w[(i & 15) * pitch + (i >> 4)] = __ldg(w + i);
}
__device__ __forceinline__ uint4 ldz16(const __nv_bfloat16* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const uint4*>(p)) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ float dot8(const uint4& u, const float* wrow) {
  const float4 wa = *reinterpret_cast<const float4*>(wrow);
  const float4 wb = *reinterpret_cast<const float4*>(wrow + 4);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  return a.x * wa.x + a.y * wa.y + b.x * wa.z + b.y * wa.w + c.x * wb.x + c.y * wb.y + d.x * wb.z + d.y * wb.w;
}

// out[n, ho, wo] = bias + sum_{r,s,c} x[n, 2ho+r-1, 2wo+s-1, c] * w[c][r][s]   (w: PyTorch [1,C,4,4])
// One warp per output pixel.  The four pixels of a filter row are contiguous in NHWC, so a lane
// reads 16-byte vectors at a 512-byte lane stride: each load instruction of the warp is one
// contiguous 512-byte span.
constexpr int kClsWarps = 8;
__global__ void __launch_bounds__(kClsWarps * 32)
classifier_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                      const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int Ho, int Wo) {
  extern __shared__ __align__(16) float s_w[];  // [16][C + 4]
  stage_filter_t(w, C, s_w);
  __syncthreads();
  const int pitch = C + 4;
  const int lane = threadIdx.x & 31;
  const int total = N * Ho * Wo;
  for (int pix = blockIdx.x * kClsWarps + (threadIdx.x >> 5); pix < total; pix += gridDim.x * kClsWarps) {
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, n = pix / (Wo * Ho);
    float acc = 0.f;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {   // two batches of 8 taps: 8 loads in flight per lane
        uint4 raw[8];
#pragma unroll
        for (int t8 = 0; t8 < 8; ++t8) {
          const int t = half * 8 + t8;
          const int h = ho * 2 + (t >> 2) - 1, ww = wo * 2 + (t & 3) - 1;
          const bool ok = h >= 0 && h < H && ww >= 0 && ww < W;
          raw[t8] = ldz16(x + (((int64_t)n * H + (ok ? h : 0)) * W + (ok ? ww : 0)) * x_ld + c0, ok);
        }
#pragma unroll
        for (int t8 = 0; t8 < 8; ++t8) acc += dot8(raw[t8], s_w + (half * 8 + t8) * pitch + c0);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pix] = acc + bias[0];
  }
}

// dx[n, h, w, c] = sum_{r,s} dout[n, (h+1-r)/2, (w+1-s)/2] * w[c][r][s]
// Work item = a 2x2 block of dx pixels x 8 channels: the block uses every one of the 16 taps once
// per dout neighbour (r = ph + 1 - 2 di, s = pw + 1 - 2 dj, di, dj in {-1, 0, 1}), so the control
// flow is uniform; nine dout scalars are loaded up front and the result leaves as 16-byte stores.
__global__ void __launch_bounds__(256)
classifier_dgrad_kernel(const float* __restrict__ dout, int N, int H, int W, int C, int Ho, int Wo,
                        const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int dx_ld) {
  extern __shared__ __align__(16) float s_w[];  // [16][C + 4]
  stage_filter_t(w, C, s_w);
  __syncthreads();
  const int pitch = C + 4;
  const int groups = C >> 3;
  const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;
  const int64_t items = (int64_t)N * Hb * Wb * groups;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(it % groups);
    const int64_t blk = it / groups;
    const int j = (int)(blk % Wb), i = (int)((blk / Wb) % Hb), n = (int)(blk / ((int64_t)Wb * Hb));
    float d[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int ho = i - 1 + a, wo = j - 1 + b;
        const bool ok = ho >= 0 && ho < Ho && wo >= 0 && wo < Wo;
        d[a][b] = ok ? __ldg(dout + ((int64_t)n * Ho + ho) * Wo + wo) : 0.f;
      }
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        float acc[8] = {0};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const int r = ph + 1 - 2 * (a - 1);
          if (r < 0 || r >= 4) continue;
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const int sx = pw + 1 - 2 * (b - 1);
            if (sx < 0 || sx >= 4) continue;
            const float* wr = s_w + (r * 4 + sx) * pitch + g * 8;
            const float4 wa = *reinterpret_cast<const float4*>(wr);
            const float4 wb = *reinterpret_cast<const float4*>(wr + 4);
            const float dv = d[a][b];
            acc[0] += dv * wa.x; acc[1] += dv * wa.y; acc[2] += dv * wa.z; acc[3] += dv * wa.w;
            acc[4] += dv * wb.x; acc[5] += dv * wb.y; acc[6] += dv * wb.z; acc[7] += dv * wb.w;
          }
        }
        const int h = 2 * i + ph, ww = 2 * j + pw;
        if (h < H && ww < W)
          *reinterpret_cast<uint4*>(dx + (((int64_t)n * H + h) * W + ww) * dx_ld + g * 8) =
              make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                         pack_bf16(acc[6], acc[7]));
      }
    }
  }
}

// dw[c][r][s] += sum_pixels dout * x ; dbias += sum dout.
// grid: x = chunk of output pixels, y = filter row r.  Block = (C/8 channel groups) x pixel lanes;
// a thread keeps the 4 taps of its row x 8 channels and walks its pixels two at a time with the
// eight 16-byte loads issued first.
__global__ void __launch_bounds__(256)
classifier_wgrad_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ x, int x_ld,
                        int N, int H, int W, int C, int Ho, int Wo, float* __restrict__ dw,
                        float* __restrict__ dbias) {
  extern __shared__ __align__(16) float s_acc[];  // [4][C]
  const int groups = C >> 3;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int r = blockIdx.y;
  const int npix = N * Ho * Wo;
  const int per = (npix + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float acc[4][8];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[s][c] = 0.f;
  float db = 0.f;
  if (pl < lanes) {
    for (int pb = p0 + pl * 2; pb < p1; pb += lanes * 2) {
      uint4 raw[2][4];
      float d[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int p = pb + u;
        const bool live = p < p1;
        const int pc = live ? p : p0;
        const int wo = pc % Wo, ho = (pc / Wo) % Ho, n = pc / (Wo * Ho);
        d[u] = live ? __ldg(dout + pc) : 0.f;
        const int h = ho * 2 + r - 1;
        const bool hok = live && h >= 0 && h < H;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int ww = wo * 2 + s - 1;
          const bool ok = hok && ww >= 0 && ww < W;
          raw[u][s] = ldz16(x + (((int64_t)n * H + (ok ? h : 0)) * W + (ok ? ww : 0)) * x_ld + g * 8, ok);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (g == 0) db += d[u];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float2 a = unpack_bf16(raw[u][s].x), b = unpack_bf16(raw[u][s].y), c = unpack_bf16(raw[u][s].z),
                       e = unpack_bf16(raw[u][s].w);
          acc[s][0] += d[u] * a.x; acc[s][1] += d[u] * a.y; acc[s][2] += d[u] * b.x; acc[s][3] += d[u] * b.y;
          acc[s][4] += d[u] * c.x; acc[s][5] += d[u] * c.y; acc[s][6] += d[u] * e.x; acc[s][7] += d[u] * e.y;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int c = 0; c < 8; ++c) atomicAdd(&s_acc[s * C + g * 8 + c], acc[s][c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    const int s = i / C, c = i - s * C;
    atomicAdd(&dw[c * 16 + r * 4 + s], s_acc[i]);
  }
  if (r == 0 && dbias != nullptr && pl < lanes && g == 0) atomicAdd(dbias, db);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_classifier_fwd(const void* x, int x_ld, int N, int H, int W, int C, const float* w,
                        const float* bias, float* out, cudaStream_t stream) {
  if (C % 8 || C > 1024) return set_error(B200_EINVAL, "classifier: C=%d must be a multiple of 8 (<= 1024)", C);
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  const int warps = N * Ho * Wo;
  int blocks = (warps + kClsWarps - 1) / kClsWarps;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)16 * (C + 4) * sizeof(float);
  static int optin = 0;
  if (!optin) {
    cudaFuncSetAttribute(classifier_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(classifier_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    optin = 1;
  }
  classifier_fwd_kernel<<<blocks, kClsWarps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ld, N, H, W, C, w, bias, out, Ho, Wo);
  return check_launch("classifier_fwd");
}

int b200_classifier_dgrad(const float* dout, int N, int H, int W, int C, const float* w, void* dx,
                          int dx_ld, cudaStream_t stream) {
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  if (C % 8 || C > 1024) return set_error(B200_EINVAL, "classifier dgrad: C=%d must be a multiple of 8 (<= 1024)", C);
  const int64_t items = (int64_t)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  int64_t blocks = (items + 255) / 256;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)16 * (C + 4) * sizeof(float);
  static int optin = 0;
  if (!optin) {
    cudaFuncSetAttribute(classifier_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    optin = 1;
  }
  classifier_dgrad_kernel<<<(int)blocks, 256, smem, stream>>>(dout, N, H, W, C, Ho, Wo, w, static_cast<__nv_bfloat16*>(dx), dx_ld);
  return check_launch("classifier_dgrad");
}

int b200_classifier_wgrad(const float* dout, const void* x, int x_ld, int N, int H, int W, int C,
                          float* dw, float* dbias, cudaStream_t stream) {
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  if (C % 8 || C > 2048) return set_error(B200_EINVAL, "classifier wgrad: C=%d must be a multiple of 8 (<= 2048)", C);
  const int groups = C / 8;
  int threads = (256 / groups) * groups;
  if (threads == 0) threads = groups;
  const int lanes = threads / groups;
  int chunks = (N * Ho * Wo + lanes * 8 - 1) / (lanes * 8);
  if (chunks > 74) chunks = 74;
  if (chunks < 1) chunks = 1;
  classifier_wgrad_kernel<<<dim3(chunks, 4), threads, (size_t)4 * C * sizeof(float), stream>>>(
      dout, static_cast<const __nv_bfloat16*>(x), x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  return check_launch("classifier_wgrad");
}

}  // extern "C"
