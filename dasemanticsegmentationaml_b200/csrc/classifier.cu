// The discriminators' last layer: Conv2d(512, 1, kernel 4, stride 2, pad 1) with bias — an
// N = 1 "GEMV" convolution (0.07 GFLOP at batch 8), run as one warp per output pixel.
//   reference: self.classifier in model/discriminator.py:13,47,97 (forward at 26,72,132)
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

// out[n, ho, wo] = bias + sum_{r,s,c} x[n, 2ho+r-1, 2wo+s-1, c] * w[c][r][s]   (w: PyTorch [1,C,4,4])
// One warp per output pixel; the filter is staged transposed ([tap][c]) in shared memory so that a
// lane's 16 channels are four conflict-free float4 reads per tap and the activations two 16-byte
// loads; the 16 taps are independent (unrolled) loads.
constexpr int kClsWarps = 8;
__global__ void __launch_bounds__(kClsWarps * 32)
classifier_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                      const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int Ho, int Wo) {
  extern __shared__ float s_w[];  // [16][C]
  for (int i = threadIdx.x; i < 16 * C; i += blockDim.x) {
    const int c = i >> 4, t = i & 15;
    s_w[t * C + c] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int total = N * Ho * Wo;
  for (int pix = blockIdx.x * kClsWarps + (threadIdx.x >> 5); pix < total; pix += gridDim.x * kClsWarps) {
    const int wo = pix % Wo, ho = (pix / Wo) % Ho, n = pix / (Wo * Ho);
    float acc = 0.f;
    for (int c0 = lane * 16; c0 < C; c0 += 512) {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int h = ho * 2 + (t >> 2) - 1, ww = wo * 2 + (t & 3) - 1;
        if (h < 0 || h >= H || ww < 0 || ww >= W) continue;
        const uint4* px = reinterpret_cast<const uint4*>(x + (((int64_t)n * H + h) * W + ww) * x_ld + c0);
        const uint4 u0 = px[0], u1 = px[1];
        const float4* wv = reinterpret_cast<const float4*>(s_w + t * C + c0);
        const float4 w0 = wv[0], w1 = wv[1], w2 = wv[2], w3 = wv[3];
        float2 f;
        f = unpack_bf16(u0.x); acc += f.x * w0.x + f.y * w0.y;
        f = unpack_bf16(u0.y); acc += f.x * w0.z + f.y * w0.w;
        f = unpack_bf16(u0.z); acc += f.x * w1.x + f.y * w1.y;
        f = unpack_bf16(u0.w); acc += f.x * w1.z + f.y * w1.w;
        f = unpack_bf16(u1.x); acc += f.x * w2.x + f.y * w2.y;
        f = unpack_bf16(u1.y); acc += f.x * w2.z + f.y * w2.w;
        f = unpack_bf16(u1.z); acc += f.x * w3.x + f.y * w3.y;
        f = unpack_bf16(u1.w); acc += f.x * w3.z + f.y * w3.w;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pix] = acc + bias[0];
  }
}

// dx[n, h, w, c] = sum_{r,s} dout[n, (h+1-r)/2, (w+1-s)/2] * w[c][r][s]
// A thread owns one channel pair (its 2 x 16 filter taps live in registers) and walks pixels;
// the block's threads cover all channels of a pixel -> coalesced 4-byte stores.
__global__ void __launch_bounds__(256)
classifier_dgrad_kernel(const float* __restrict__ dout, int N, int H, int W, int C, int Ho, int Wo,
                        const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int dx_ld) {
  const int pairs = C / 2;
  const int ppb = blockDim.x / pairs;             // pixels handled concurrently by a block
  const int c = (threadIdx.x % pairs) * 2;
  const int sub = threadIdx.x / pairs;
  float w0[16], w1[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    w0[t] = w[c * 16 + t];
    w1[t] = w[(c + 1) * 16 + t];
  }
  const int total = N * H * W;
  if (sub >= ppb) return;
  for (int p = blockIdx.x * ppb + sub; p < total; p += gridDim.x * ppb) {
    const int ww = p % W, h = (p / W) % H, n = p / (W * H);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int hr = h + 1 - r;
      if (hr < 0 || (hr & 1) || (hr >> 1) >= Ho) continue;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int wr = ww + 1 - s;
        if (wr < 0 || (wr & 1) || (wr >> 1) >= Wo) continue;
        const float d = __ldg(dout + ((int64_t)n * Ho + (hr >> 1)) * Wo + (wr >> 1));
        a0 += d * w0[r * 4 + s];
        a1 += d * w1[r * 4 + s];
      }
    }
    *reinterpret_cast<uint32_t*>(dx + (int64_t)p * dx_ld + c) = pack_bf16(a0, a1);
  }
}

// dw[c][r][s] += sum_pixels dout * x ; dbias += sum dout.  grid: x = (r,s) tap, y = pixel chunk;
// thread = channel pair; four independent pixels in flight per iteration.
__global__ void __launch_bounds__(256)
classifier_wgrad_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ x, int x_ld,
                        int N, int H, int W, int C, int Ho, int Wo, float* __restrict__ dw,
                        float* __restrict__ dbias) {
  const int tap = blockIdx.x, r = tap >> 2, s = tap & 3;
  const int npix = N * Ho * Wo;
  const int per = (npix + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * per, p1 = min(npix, p0 + per);
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {
    float a0 = 0.f, a1 = 0.f;
    for (int pb = p0; pb < p1; pb += 4) {
      float d[4];
      uint32_t v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = pb + u;
        d[u] = 0.f;
        v[u] = 0u;
        if (p < p1) {
          const int wo = p % Wo, ho = (p / Wo) % Ho, n = p / (Wo * Ho);
          const int h = ho * 2 + r - 1, ww = wo * 2 + s - 1;
          if (h >= 0 && h < H && ww >= 0 && ww < W) {
            d[u] = __ldg(dout + p);
            v[u] = *reinterpret_cast<const uint32_t*>(x + (((int64_t)n * H + h) * W + ww) * x_ld + c);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = unpack_bf16(v[u]);
        a0 += d[u] * f.x;
        a1 += d[u] * f.y;
      }
    }
    atomicAdd(&dw[c * 16 + tap], a0);
    atomicAdd(&dw[(c + 1) * 16 + tap], a1);
  }
  if (tap == 0 && dbias != nullptr && threadIdx.x == 0) {
    float db = 0.f;
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
  if (C % 512 && C != 64 && C != 128 && C != 256) return set_error(B200_EINVAL, "classifier: C=%d unsupported", C);
  const int warps = N * Ho * Wo;
  int blocks = (warps + kClsWarps * 2 - 1) / (kClsWarps * 2);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  static int optin = 0;
  if (!optin) {
    cudaFuncSetAttribute(classifier_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    optin = 1;
  }
  classifier_fwd_kernel<<<blocks, kClsWarps * 32, (size_t)16 * C * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(x), x_ld, N, H, W, C, w, bias, out, Ho, Wo);
  return check_launch("classifier_fwd");
}

int b200_classifier_dgrad(const float* dout, int N, int H, int W, int C, const float* w, void* dx,
                          int dx_ld, cudaStream_t stream) {
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  if (C % 2 || C / 2 > 256) return set_error(B200_EINVAL, "classifier dgrad: C=%d unsupported", C);
  const int ppb = 256 / (C / 2);
  int64_t blocks = ((int64_t)N * H * W + ppb * 8 - 1) / (ppb * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  classifier_dgrad_kernel<<<(int)blocks, 256, 0, stream>>>(dout, N, H, W, C, Ho, Wo, w, static_cast<__nv_bfloat16*>(dx), dx_ld);
  return check_launch("classifier_dgrad");
}

int b200_classifier_wgrad(const float* dout, const void* x, int x_ld, int N, int H, int W, int C,
                          float* dw, float* dbias, cudaStream_t stream) {
  const int Ho = (H + 2 - 4) / 2 + 1, Wo = (W + 2 - 4) / 2 + 1;
  int chunks = (N * Ho * Wo + 63) / 64;
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  classifier_wgrad_kernel<<<dim3(16, chunks), 256, 0, stream>>>(dout, static_cast<const __nv_bfloat16*>(x), x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  return check_launch("classifier_wgrad");
}

}  // extern "C"
