// Stem convolution: 3 -> 32 channels, 3x3, stride 2, pad 1, bias-free, on the fp32 NCHW image.
//   reference: features[0] = ConvX(3, 32, 3, 2)  (model/stdcnet.py:171, 6-15)
// K = 3*3*3 = 27 is too thin for the tap-by-tap implicit GEMM (each tap would be a K = 3 slab), and
// the CUDA-core version of the layer was 5-10x off the HBM roofline (ncu: forward 133 us, filter
// gradient 272 us at N=8, 512x1024).  Instead the image is unfolded once into bf16 rows of 32
// values per output pixel (64-byte pixels; this also fuses the NCHW->NHWC and fp32->bf16
// conversions), after which the forward is a K = 32 1x1 implicit GEMM with the BatchNorm statistics
// in its epilogue and the filter gradient a 1x1 wgrad GEMM, both on the tensor pipe.
// Traffic: image 12 B per input pixel in, 64 B per output pixel out (HBM-bound).
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

// One thread = one output pixel: 27 loads (a warp reads 9 contiguous ~260-byte spans per plane row)
// and four 16-byte stores (a warp writes 2 KB contiguous).
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, int N, int H, int W, __nv_bfloat16* __restrict__ col,
                   int col_ld, int Ho, int Wo) {
  const int64_t npix = (int64_t)N * Ho * Wo;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int wo = (int)(p % Wo);
    const int64_t t2 = p / Wo;
    const int ho = (int)(t2 % Ho), n = (int)(t2 / Ho);
    float v[32];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + ((int64_t)n * 3 + ci) * H * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = ho * 2 + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int w = wo * 2 + s - 1;
          float x = 0.f;
          if (h >= 0 && h < H && w >= 0 && w < W) x = __ldg(plane + (int64_t)h * W + w);
          v[ci * 9 + r * 3 + s] = x;
        }
      }
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
    uint4* o = reinterpret_cast<uint4*>(col + p * col_ld);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                        pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_stem_im2col(const float* img, int N, int H, int W, void* col, int col_ld, cudaStream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (col_ld < 32 || col_ld % 8) return set_error(B200_EINVAL, "stem_im2col: pixel stride %d must be >= 32 and a multiple of 8", col_ld);
  int64_t blocks = ((int64_t)N * Ho * Wo + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  stem_im2col_kernel<<<(int)blocks, 256, 0, stream>>>(img, N, H, W, static_cast<__nv_bfloat16*>(col), col_ld, Ho, Wo);
  return check_launch("stem_im2col");
}

}  // extern "C"
