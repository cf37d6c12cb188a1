// The discriminators' last layer: Conv2d(512, 1, kernel 4, stride 2, pad 1) with bias — an
// N = 1 "GEMV" convolution (0.07 GFLOP at batch 8), run as one warp per output pixel.
//   reference: self.classifier in model/discriminator.py:13,47,97 (forward at 26,72,132)
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

// out[n, ho, wo] = bias + sum_{r,s,c} x[n, 2ho+r-1, 2wo+s-1, c] * w[c][r][s]   (w: PyTorch [1,C,4,4])
__global__ void __launch_bounds__(256)
classifier_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                      const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int Ho, int Wo) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N * Ho * Wo) return;
  const int wo = warp % Wo, ho = (warp / Wo) % Ho, n = warp / (Wo * Ho);
  float acc = 0.f;
  for (int r = 0; r < 4; ++r) {
    const int h = ho * 2 + r - 1;
    if (h < 0 || h >= H) continue;
    for (int s = 0; s < 4; ++s) {
      const int ww = wo * 2 + s - 1;
      if (ww < 0 || ww >= W) continue;
      const __nv_bfloat16* px = x + (((int64_t)n * H + h) * W + ww) * x_ld;
      for (int c = lane * 2; c < C; c += 64) {
        const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(px + c));
        acc += v.x * __ldg(w + c * 16 + r * 4 + s) + v.y * __ldg(w + (c + 1) * 16 + r * 4 + s);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[warp] = acc + bias[0];
}

// dx[n, h, w, c] = sum_{r,s} dout[n, (h+1-r)/2, (w+1-s)/2] * w[c][r][s]
__global__ void __launch_bounds__(256)
classifier_dgrad_kernel(const float* __restrict__ dout, int N, int H, int W, int C, int Ho, int Wo,
                        const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int dx_ld) {
  const int64_t total = (int64_t)N * H * W * (C / 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C / 2)) * 2;
    const int64_t p = i / (C / 2);
    const int ww = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int hr = h + 1 - r;
      if (hr < 0 || (hr & 1) || (hr >> 1) >= Ho) continue;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int wr = ww + 1 - s;
        if (wr < 0 || (wr & 1) || (wr >> 1) >= Wo) continue;
        const float d = dout[((int64_t)n * Ho + (hr >> 1)) * Wo + (wr >> 1)];
        a0 += d * __ldg(w + c * 16 + r * 4 + s);
        a1 += d * __ldg(w + (c + 1) * 16 + r * 4 + s);
      }
    }
    *reinterpret_cast<uint32_t*>(dx + p * dx_ld + c) = pack_bf16(a0, a1);
  }
}

// dw[c][r][s] += sum_pixels dout * x ; dbias += sum dout.  grid: x = (r,s) tap, y = pixel chunk;
// thread = channel pair.
__global__ void __launch_bounds__(256)
classifier_wgrad_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ x, int x_ld,
                        int N, int H, int W, int C, int Ho, int Wo, float* __restrict__ dw,
                        float* __restrict__ dbias) {
  const int tap = blockIdx.x, r = tap >> 2, s = tap & 3;
  const int npix = N * Ho * Wo;
  const int per = (npix + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * per, p1 = min(npix, p0 + per);
  float db = 0.f;
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {
    float a0 = 0.f, a1 = 0.f;
    for (int p = p0; p < p1; ++p) {
      const int wo = p % Wo, ho = (p / Wo) % Ho, n = p / (Wo * Ho);
      const int h = ho * 2 + r - 1, ww = wo * 2 + s - 1;
      if (h < 0 || h >= H || ww < 0 || ww >= W) continue;
      const float d = dout[p];
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (((int64_t)n * H + h) * W + ww) * x_ld + c));
      a0 += d * v.x;
      a1 += d * v.y;
    }
    atomicAdd(&dw[c * 16 + tap], a0);
    atomicAdd(&dw[(c + 1) * 16 + tap], a1);
  }
  if (tap == 0 && dbias != nullptr && threadIdx.x == 0) {
    for (int p = p0; p < p1; ++p) db += dout[p];
    atomicAdd(dbias, db);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_classifier_fwd(const void* x, int x_ld, int N, int H, int W, int C, const float* w,
                        const float* bias, float* out, cudaStream_t stream) {
  if (C % 64) return set_error(B200_EINVAL, "classifier: C=%d must be a multiple of 64", C);
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  const int warps = N * Ho * Wo;
  classifier_fwd_kernel<<<(warps + 7) / 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ld, N, H, W, C, w, bias, out, Ho, Wo);
  return check_launch("classifier_fwd");
}

int b200_classifier_dgrad(const float* dout, int N, int H, int W, int C, const float* w, void* dx,
                          int dx_ld, cudaStream_t stream) {
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  int64_t total = (int64_t)N * H * W * (C / 2);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  classifier_dgrad_kernel<<<(int)blocks, 256, 0, stream>>>(dout, N, H, W, C, Ho, Wo, w, static_cast<__nv_bfloat16*>(dx), dx_ld);
  return check_launch("classifier_dgrad");
}

int b200_classifier_wgrad(const float* dout, const void* x, int x_ld, int N, int H, int W, int C,
                          float* dw, float* dbias, cudaStream_t stream) {
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  int chunks = (N * Ho * Wo + 255) / 256;
  if (chunks > 32) chunks = 32;
  if (chunks < 1) chunks = 1;
  classifier_wgrad_kernel<<<dim3(16, chunks), 256, 0, stream>>>(dout, static_cast<const __nv_bfloat16*>(x), x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  return check_launch("classifier_wgrad");
}

}  // extern "C"
