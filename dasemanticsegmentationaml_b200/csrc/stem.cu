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

// Both kernels work on a tile of TSH x TSW output pixels whose (2*TSH+1) x (2*TSW+1) x 3 input
// patch is staged in shared memory with coalesced reads of the NCHW planes (bf16-rounded, so the
// layer computes exactly what a bf16 tensor-core conv would).
constexpr int TSH = 8, TSW = 32;                      // forward tile
constexpr int PH = 2 * TSH + 1, PW = 2 * TSW + 1;

template <int TH_, int TW_>
__device__ __forceinline__ void stage_patch(const float* __restrict__ img, int n, int H, int W, int h0,
                                            int w0, float* s_patch) {
  constexpr int ph = 2 * TH_ + 1, pw = 2 * TW_ + 1;
  for (int i = threadIdx.x; i < 3 * ph * pw; i += blockDim.x) {
    const int c = i / (ph * pw), rem = i - c * (ph * pw);
    const int y = rem / pw, xx = rem - y * pw;
    const int h = 2 * h0 - 1 + y, w = 2 * w0 - 1 + xx;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W)
      v = __bfloat162float(__float2bfloat16(__ldg(img + ((size_t)(n * 3 + c) * H + h) * W + w)));
    s_patch[i] = v;
  }
}

// Forward: one thread = one output pixel x all 32 channels.  The filter is read from shared
// memory with warp-uniform (broadcast) float4 loads -- one wavefront per 4 weights -- and a warp
// writes 32 consecutive 64-byte pixels, i.e. 2 KB contiguous.
__global__ void __launch_bounds__(256, 2)
stem_fwd_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ w,
                __nv_bfloat16* __restrict__ z, int z_ld, int Ho, int Wo, float* __restrict__ stats) {
  __shared__ __align__(16) float s_w[27][kStemCout];  // [ci*9 + r*3 + s][co], bf16-rounded
  __shared__ float s_acc[2][kStemCout];
  for (int i = threadIdx.x; i < 27 * kStemCout; i += blockDim.x) {
    const int co = i % kStemCout, t = i / kStemCout;
    s_w[t][co] = __bfloat162float(__float2bfloat16(w[co * 27 + t]));
  }
  if (threadIdx.x < 2 * kStemCout) (&s_acc[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  float s1[kStemCout], s2[kStemCout];
#pragma unroll
  for (int c = 0; c < kStemCout; ++c) s1[c] = s2[c] = 0.f;
  const int npix = N * Ho * Wo;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int wo = p % Wo;
    const int t2 = p / Wo;
    const int ho = t2 % Ho;
    const int n = t2 / Ho;
    float acc[kStemCout];
#pragma unroll
    for (int c = 0; c < kStemCout; ++c) acc[c] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + (size_t)(n * 3 + ci) * H * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = ho * 2 + r - 1;
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          const int ww = wo * 2 + sx - 1;
          float v = 0.f;
          if (h >= 0 && h < H && ww >= 0 && ww < W)
            v = __bfloat162float(__float2bfloat16(__ldg(plane + (size_t)h * W + ww)));
          const float4* wr = reinterpret_cast<const float4*>(&s_w[ci * 9 + r * 3 + sx][0]);
#pragma unroll
          for (int q = 0; q < kStemCout / 4; ++q) {
            const float4 w4 = wr[q];
            acc[4 * q] += v * w4.x;
            acc[4 * q + 1] += v * w4.y;
            acc[4 * q + 2] += v * w4.z;
            acc[4 * q + 3] += v * w4.w;
          }
        }
      }
    }
    uint4* o = reinterpret_cast<uint4*>(z + (size_t)p * z_ld);
#pragma unroll
    for (int q = 0; q < kStemCout / 8; ++q)
      o[q] = make_uint4(pack_bf16(acc[8 * q], acc[8 * q + 1]), pack_bf16(acc[8 * q + 2], acc[8 * q + 3]),
                        pack_bf16(acc[8 * q + 4], acc[8 * q + 5]), pack_bf16(acc[8 * q + 6], acc[8 * q + 7]));
    if (stats != nullptr) {
#pragma unroll
      for (int c = 0; c < kStemCout; ++c) {
        const float rr = __bfloat162float(__float2bfloat16(acc[c]));
        s1[c] += rr;
        s2[c] += rr * rr;
      }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int c = 0; c < kStemCout; ++c) {
      const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_acc[0][c], a);
        atomicAdd(&s_acc[1][c], b);
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kStemCout)
      atomicAdd(&stats[threadIdx.x], (&s_acc[0][0])[threadIdx.x]);
  }
}

// dw[co][ci][r][s] += sum_pixels dz[pixel][co] * img[ci][2*ho + r - 1][2*wo + s - 1]
// A [32 x 27] = dz^T [32 x P] * xcol [P x 27] product on the CUDA cores: a block stages the input
// patch and the dz tile of 8 x 16 pixels, then 288 threads = 4 pixel groups x (8 co-quads x 9
// tap-triples) accumulate 4x3 register tiles; a block walks several tiles before its totals go to
// the global gradient with one atomic per filter element.
constexpr int WSH = 8, WSW = 16;
constexpr int WPH = 2 * WSH + 1, WPW = 2 * WSW + 1;
constexpr int kSwThreads = 288;
__global__ void __launch_bounds__(kSwThreads)
stem_wgrad_kernel(const float* __restrict__ img, int N, int H, int W,
                  const __nv_bfloat16* __restrict__ dz, int dz_ld, int Ho, int Wo,
                  float* __restrict__ dw) {
  __shared__ float s_dz[WSH * WSW][kStemCout + 1];
  __shared__ float s_patch[3 * WPH * WPW];
  __shared__ float s_acc[kStemCout * 27];
  for (int i = threadIdx.x; i < kStemCout * 27; i += kSwThreads) s_acc[i] = 0.f;
  const int grp = threadIdx.x / 72;          // pixel group 0..3
  const int lt = threadIdx.x % 72;
  const int cq = lt & 7;                     // co quad: co = cq*4 .. +3
  const int tt = lt >> 3;                    // tap triple: t = tt*3 .. +2  (= ci*3 + r, s = 0..2)
  const int ci = tt / 3, r = tt % 3;
  float acc[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) acc[a][b] = 0.f;
  const int tiles_w = (Wo + WSW - 1) / WSW, tiles_h = (Ho + WSH - 1) / WSH;
  const int total = tiles_w * tiles_h * N;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / (tiles_w * tiles_h);
    const int t_in = tile - n * tiles_w * tiles_h;
    const int h0 = (t_in / tiles_w) * WSH, w0 = (t_in % tiles_w) * WSW;
    __syncthreads();
    stage_patch<WSH, WSW>(img, n, H, W, h0, w0, s_patch);
    for (int i = threadIdx.x; i < WSH * WSW * (kStemCout / 2); i += kSwThreads) {
      const int px = i / (kStemCout / 2), c2 = i % (kStemCout / 2);
      const int ho = h0 + px / WSW, wo = w0 + px % WSW;
      float2 v = make_float2(0.f, 0.f);
      if (ho < Ho && wo < Wo)
        v = unpack_bf16(*reinterpret_cast<const uint32_t*>(dz + (((size_t)n * Ho + ho) * Wo + wo) * dz_ld + c2 * 2));
      s_dz[px][c2 * 2] = v.x;
      s_dz[px][c2 * 2 + 1] = v.y;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = grp; px < WSH * WSW; px += 4) {
      const int hl = px / WSW, wl = px % WSW;
      float d[4], x[3];
#pragma unroll
      for (int a = 0; a < 4; ++a) d[a] = s_dz[px][cq * 4 + a];
      const float* row = &s_patch[(ci * WPH + 2 * hl + r) * WPW + 2 * wl];
#pragma unroll
      for (int b = 0; b < 3; ++b) x[b] = row[b];
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
  if ((int64_t)N * H * W >= (1ll << 31)) return set_error(B200_EINVAL, "stem: tensor too large");
  int64_t blocks = ((int64_t)N * Ho * Wo + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  stem_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(img, N, H, W, w, static_cast<__nv_bfloat16*>(z), z_ld, Ho, Wo, stats);
  return check_launch("stem_fwd");
}

int b200_stem_wgrad(const float* img, int N, int H, int W, const void* dz, int dz_ld, float* dw,
                    cudaStream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int tiles = ((Wo + WSW - 1) / WSW) * ((Ho + WSH - 1) / WSH) * N;
  int blocks = tiles < 148 * 4 ? tiles : 148 * 4;
  if (blocks < 1) blocks = 1;
  stem_wgrad_kernel<<<blocks, kSwThreads, 0, stream>>>(img, N, H, W, static_cast<const __nv_bfloat16*>(dz), dz_ld, Ho, Wo, dw);
  return check_launch("stem_wgrad");
}

}  // extern "C"
