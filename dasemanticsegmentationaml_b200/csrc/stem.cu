// Stem convolution: 3 -> 32 channels, 3x3, stride 2, pad 1, bias-free, reading the fp32 NCHW image
// directly (fuses the layout + precision conversion) and writing bf16 NHWC.  K = 27 is far too
// thin for the tensor pipe; the layer is HBM-bound (50 MB in, 67 MB out at N=8, 512x1024).
//   reference: features[0] = ConvX(3, 32, 3, 2)  (model/stdcnet.py:171, 6-15)
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

constexpr int kStemCout = 32;

// One thread = one output pixel x 8 output channels (4 threads per pixel).  Inputs are rounded to
// bf16 before use so that the layer computes exactly what a bf16 tensor-core conv would.
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ w,
                __nv_bfloat16* __restrict__ z, int z_ld, int Ho, int Wo, float* __restrict__ stats) {
  __shared__ float s_w[27][kStemCout];  // [ci*9 + r*3 + s][co], bf16-rounded
  __shared__ float s_acc[2][kStemCout];
  for (int i = threadIdx.x; i < 27 * kStemCout; i += blockDim.x) {
    const int co = i % kStemCout, t = i / kStemCout;
    s_w[t][co] = __bfloat162float(__float2bfloat16(w[co * 27 + t]));
  }
  if (threadIdx.x < 2 * kStemCout) (&s_acc[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const int q = threadIdx.x & 3;  // channel octet
  float s1[8] = {0}, s2[8] = {0};
  const int64_t npix = (int64_t)N * Ho * Wo;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; p < npix;
       p += ((int64_t)gridDim.x * blockDim.x) >> 2) {
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int n = (int)(p / ((int64_t)Wo * Ho));
    float acc[8] = {0};
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + ((int64_t)n * 3 + ci) * H * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = ho * 2 + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ww = wo * 2 + s - 1;
          float v = 0.f;
          if (h >= 0 && h < H && ww >= 0 && ww < W)
            v = __bfloat162float(__float2bfloat16(__ldg(plane + (int64_t)h * W + ww)));
          const float4 wa = *reinterpret_cast<const float4*>(&s_w[ci * 9 + r * 3 + s][q * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&s_w[ci * 9 + r * 3 + s][q * 8 + 4]);
          acc[0] += v * wa.x; acc[1] += v * wa.y; acc[2] += v * wa.z; acc[3] += v * wa.w;
          acc[4] += v * wb.x; acc[5] += v * wb.y; acc[6] += v * wb.z; acc[7] += v * wb.w;
        }
      }
    }
    *reinterpret_cast<uint4*>(z + p * z_ld + q * 8) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                   pack_bf16(acc[6], acc[7]));
    if (stats != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float r = __bfloat162float(__float2bfloat16(acc[j]));
        s1[j] += r;
        s2[j] += r * r;
      }
    }
  }
  if (stats != nullptr) {
    // lanes with equal (lane & 3) hold the same channel octet
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
      }
    }
    if ((threadIdx.x & 31) < 4) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_acc[0][q * 8 + j], s1[j]);
        atomicAdd(&s_acc[1][q * 8 + j], s2[j]);
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kStemCout)
      atomicAdd(&stats[threadIdx.x], (&s_acc[0][0])[threadIdx.x]);
  }
}

// dw[co][ci][r][s] += sum_pixels dz[pixel][co] * img[ci][2*ho + r - 1][2*wo + s - 1]
// One thread = one pixel x 4 output channels (8 threads per pixel), 108 accumulators, grid-stride
// over pixels, block-level reduction in shared memory, then one atomic per filter element.
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ img, int N, int H, int W,
                  const __nv_bfloat16* __restrict__ dz, int dz_ld, int Ho, int Wo,
                  float* __restrict__ dw) {
  __shared__ float s_acc[kStemCout * 27];
  for (int i = threadIdx.x; i < kStemCout * 27; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int q = threadIdx.x & 7;  // channel quad
  float acc[27][4];
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  const int64_t npix = (int64_t)N * Ho * Wo;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; p < npix;
       p += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int n = (int)(p / ((int64_t)Wo * Ho));
    const uint2 u = *reinterpret_cast<const uint2*>(dz + p * dz_ld + q * 4);
    const float2 d01 = unpack_bf16(u.x), d23 = unpack_bf16(u.y);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + ((int64_t)n * 3 + ci) * H * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = ho * 2 + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ww = wo * 2 + s - 1;
          float v = 0.f;
          if (h >= 0 && h < H && ww >= 0 && ww < W)
            v = __bfloat162float(__float2bfloat16(__ldg(plane + (int64_t)h * W + ww)));
          const int t = ci * 9 + r * 3 + s;
          acc[t][0] += v * d01.x;
          acc[t][1] += v * d01.y;
          acc[t][2] += v * d23.x;
          acc[t][3] += v * d23.y;
        }
      }
    }
  }
  // lanes with equal (lane & 7) hold the same channel quad: reduce over xor 8, 16
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[t][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&s_acc[(q * 4 + j) * 27 + t], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < kStemCout * 27; i += blockDim.x) atomicAdd(&dw[i], s_acc[i]);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_stem_fwd(const float* img, int N, int H, int W, const float* w, void* z, int z_ld,
                  float* stats, cudaStream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t threads = (int64_t)N * Ho * Wo * 4;
  int64_t blocks = (threads + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  stem_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(img, N, H, W, w, static_cast<__nv_bfloat16*>(z), z_ld, Ho, Wo, stats);
  return check_launch("stem_fwd");
}

int b200_stem_wgrad(const float* img, int N, int H, int W, const void* dz, int dz_ld, float* dw,
                    cudaStream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t threads = (int64_t)N * Ho * Wo * 8;
  int64_t blocks = (threads + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  stem_wgrad_kernel<<<(int)blocks, 256, 0, stream>>>(img, N, H, W, static_cast<const __nv_bfloat16*>(dz), dz_ld, Ho, Wo, dw);
  return check_launch("stem_wgrad");
}

}  // extern "C"
