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
// A [32 x 27] = dz^T [32 x P] * xcol [P x 27] product on the CUDA cores, tiled through shared
// memory: a block stages 64 pixels (dz as fp32, the 27-tap input column as fp32), then 288 threads
// = 4 pixel groups x (8 co-quads x 9 tap-triples) accumulate 4x3 register tiles; block totals go
// to the global gradient with one atomic per filter element.
constexpr int kSwP = 64;       // pixels per stage
constexpr int kSwThreads = 288;
__global__ void __launch_bounds__(kSwThreads)
stem_wgrad_kernel(const float* __restrict__ img, int N, int H, int W,
                  const __nv_bfloat16* __restrict__ dz, int dz_ld, int Ho, int Wo,
                  float* __restrict__ dw) {
  __shared__ float s_dz[kSwP][kStemCout + 1];
  __shared__ float s_x[kSwP][28];
  __shared__ float s_acc[kStemCout * 27];
  for (int i = threadIdx.x; i < kStemCout * 27; i += kSwThreads) s_acc[i] = 0.f;
  const int grp = threadIdx.x / 72;          // pixel group 0..3
  const int lt = threadIdx.x % 72;
  const int cq = lt & 7;                     // co quad: co = cq*4 .. +3
  const int tt = lt >> 3;                    // tap triple: t = tt*3 .. +2
  float acc[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) acc[a][b] = 0.f;
  const int npix = N * Ho * Wo;
  for (int base = blockIdx.x * kSwP; base < npix; base += gridDim.x * kSwP) {
    __syncthreads();
    for (int i = threadIdx.x; i < kSwP * kStemCout / 2; i += kSwThreads) {
      const int px = i / (kStemCout / 2), c2 = i % (kStemCout / 2);
      float2 v = make_float2(0.f, 0.f);
      if (base + px < npix)
        v = unpack_bf16(*reinterpret_cast<const uint32_t*>(dz + (size_t)(base + px) * dz_ld + c2 * 2));
      s_dz[px][c2 * 2] = v.x;
      s_dz[px][c2 * 2 + 1] = v.y;
    }
    for (int i = threadIdx.x; i < kSwP * 27; i += kSwThreads) {
      const int px = i / 27, t = i % 27;
      float v = 0.f;
      const int p = base + px;
      if (p < npix) {
        const int wo = p % Wo, ho = (p / Wo) % Ho, n = p / (Wo * Ho);
        const int ci = t / 9, r = (t % 9) / 3, sx = t % 3;
        const int h = ho * 2 + r - 1, ww = wo * 2 + sx - 1;
        if (h >= 0 && h < H && ww >= 0 && ww < W)
          v = __bfloat162float(__float2bfloat16(__ldg(img + ((size_t)(n * 3 + ci) * H + h) * W + ww)));
      }
      s_x[px][t] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = grp; px < kSwP; px += 4) {
      float d[4], x[3];
#pragma unroll
      for (int a = 0; a < 4; ++a) d[a] = s_dz[px][cq * 4 + a];
#pragma unroll
      for (int b = 0; b < 3; ++b) x[b] = s_x[px][tt * 3 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a][b] += d[a] * x[b];
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) atomicAdd(&s_acc[(cq * 4 + a) * 27 + tt * 3 + b], acc[a][b]);
  __syncthreads();
  for (int i = threadIdx.x; i < kStemCout * 27; i += kSwThreads) atomicAdd(&dw[i], s_acc[i]);
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
  int64_t blocks = ((int64_t)N * Ho * Wo + kSwP - 1) / kSwP;
  if (blocks > 148 * 6) blocks = 148 * 6;
  if (blocks < 1) blocks = 1;
  stem_wgrad_kernel<<<(int)blocks, kSwThreads, 0, stream>>>(img, N, H, W, static_cast<const __nv_bfloat16*>(dz), dz_ld, Ho, Wo, dw);
  return check_launch("stem_wgrad");
}

}  // extern "C"
