// Depthwise stride-2 convolutions on bf16 NHWC (HBM-bound; one thread = one pixel x 8 channels,
// 16-byte accesses, consecutive threads walk the channel groups of one pixel -> coalesced).
//   reference: avd_layer dw3x3 s2 p1, bias-free (model/stdcnet.py:24-28, 73-77), the
//   AvgPool2d(3,2,1) skip of CatBottleneck (stdcnet.py:78,108-109, count_include_pad=True),
//   dw4x4 s2 p1 with bias of the depthwise-separable discriminators (model/discriminator.py:35-45).
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                            pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// Forward.  w: fp32 [C][K*K] (PyTorch [C,1,K,K]).  Writes z = dwconv(x) (+bias, +act) and, when
// `pool` is given, pool = avg_pool2d(x, 3, 2, 1) (only with K == 3).  stats (optional): [2][C]
// sum / sum-of-squares of the bf16-rounded z.
template <int K>
__global__ void __launch_bounds__(256)
dwconv_s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                     const float* __restrict__ w, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ z, int z_ld, __nv_bfloat16* __restrict__ pool,
                     int pool_ld, int Ho, int Wo, int act, float slope, float* __restrict__ stats) {
  extern __shared__ float s_acc[];  // [2][C] when stats
  const int groups = C >> 3;
  const int py = blockDim.x / groups;
  const int g = threadIdx.x % groups, ty = threadIdx.x / groups;
  if (stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  float wt[K * K][8];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wt[t][j] = w[(g * 8 + j) * K * K + t];
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = bias != nullptr ? bias[g * 8 + j] : 0.f;
  float s1[8] = {0}, s2[8] = {0};
  const int64_t npix = (int64_t)N * Ho * Wo;
  for (int64_t p = (int64_t)blockIdx.x * py + ty; p < npix; p += (int64_t)gridDim.x * py) {
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int n = (int)(p / ((int64_t)Wo * Ho));
    float acc[8], pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] = bs[j];
      pl[j] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int h = ho * 2 + r - 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int ww = wo * 2 + s - 1;
        if (ww < 0 || ww >= W) continue;
        float v[8];
        ld8(x + (((int64_t)n * H + h) * W + ww) * x_ld + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] += v[j] * wt[r * K + s][j];
          pl[j] += v[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = acc[j];
      if (act == 2) t = t > 0.f ? t : t * slope;
      acc[j] = t;
    }
    st8(z + p * z_ld + g * 8, acc);
    if (pool != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) pl[j] *= (1.f / 9.f);
      st8(pool + p * pool_ld + g * 8, pl);
    }
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
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_acc[g * 8 + j], s1[j]);
      atomicAdd(&s_acc[C + g * 8 + j], s2[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], s_acc[i]);
  }
}

// Data gradient: dx[n,h,w,c] = sum_{r,s} dz[n,(h+1-r)/2,(w+1-s)/2,c] * w[c,r,s]
//                              (+ sum over the 3x3/s2 pooling windows covering (h,w) of dpool / 9)
template <int K>
__global__ void __launch_bounds__(256)
dwconv_s2_dgrad_kernel(const __nv_bfloat16* __restrict__ dz, int dz_ld,
                       const __nv_bfloat16* __restrict__ dpool, int dpool_ld, int N, int H, int W,
                       int C, int Ho, int Wo, const float* __restrict__ w,
                       __nv_bfloat16* __restrict__ dx, int dx_ld) {
  const int groups = C >> 3;
  const int64_t total = (int64_t)N * H * W * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int64_t p = i / groups;
    const int ww = (int)(p % W);
    const int h = (int)((p / W) % H);
    const int n = (int)(p / ((int64_t)W * H));
    float acc[8] = {0};
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int hr = h + 1 - r;
      if (hr < 0 || (hr & 1)) continue;
      const int ho = hr >> 1;
      if (ho >= Ho) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int wr = ww + 1 - s;
        if (wr < 0 || (wr & 1)) continue;
        const int wo = wr >> 1;
        if (wo >= Wo) continue;
        const int64_t q = ((int64_t)n * Ho + ho) * Wo + wo;
        float v[8];
        ld8(dz + q * dz_ld + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j] * __ldg(w + (g * 8 + j) * K * K + r * K + s);
        if (dpool != nullptr) {
          float u[8];
          ld8(dpool + q * dpool_ld + g * 8, u);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += u[j] * (1.f / 9.f);
        }
      }
    }
    st8(dx + p * dx_ld + g * 8, acc);
  }
}

// Filter / bias gradient: dw[c][r][s] += sum_pixels dz[.,c] * x[.*2 + tap, c]; dbias[c] += sum dz.
template <int K>
__global__ void __launch_bounds__(256)
dwconv_s2_wgrad_kernel(const __nv_bfloat16* __restrict__ dz, int dz_ld,
                       const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                       int Ho, int Wo, float* __restrict__ dw, float* __restrict__ dbias) {
  extern __shared__ float s_acc[];  // [C][K*K + 1]
  const int groups = C >> 3;
  const int py = blockDim.x / groups;
  const int g = threadIdx.x % groups, ty = threadIdx.x / groups;
  constexpr int KK1 = K * K + 1;
  for (int i = threadIdx.x; i < C * KK1; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float acc[K * K][8];
  float ab[8] = {0};
#pragma unroll
  for (int t = 0; t < K * K; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int64_t npix = (int64_t)N * Ho * Wo;
  for (int64_t p = (int64_t)blockIdx.x * py + ty; p < npix; p += (int64_t)gridDim.x * py) {
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int n = (int)(p / ((int64_t)Wo * Ho));
    float d[8];
    ld8(dz + p * dz_ld + g * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) ab[j] += d[j];
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int h = ho * 2 + r - 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int ww = wo * 2 + s - 1;
        if (ww < 0 || ww >= W) continue;
        float v[8];
        ld8(x + (((int64_t)n * H + h) * W + ww) * x_ld + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r * K + s][j] += d[j] * v[j];
      }
    }
  }
#pragma unroll
  for (int t = 0; t < K * K; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(g * 8 + j) * KK1 + t], acc[t][j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(g * 8 + j) * KK1 + K * K], ab[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < C * KK1; i += blockDim.x) {
    const int c = i / KK1, t = i - c * KK1;
    if (t < K * K)
      atomicAdd(&dw[c * K * K + t], s_acc[i]);
    else if (dbias != nullptr)
      atomicAdd(&dbias[c], s_acc[i]);
  }
}

static int threads_for(int C) {
  const int groups = C / 8;
  int t = (256 / groups) * groups;
  return t == 0 ? groups : t;
}
static int grid_for(int64_t items, int per_block, int cap) {
  int64_t b = (items + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_dwconv_s2_fwd(const void* x, int x_ld, int N, int H, int W, int C, int K, const float* w,
                       const float* bias, void* z, int z_ld, void* pool, int pool_ld, int act,
                       float slope, float* stats, cudaStream_t stream) {
  if (C % 8 || C > 2048) return set_error(B200_EINVAL, "dwconv: C=%d must be a multiple of 8 (<= 2048)", C);
  if (K != 3 && K != 4) return set_error(B200_EINVAL, "dwconv: K=%d unsupported", K);
  if (pool != nullptr && K != 3) return set_error(B200_EINVAL, "dwconv: fused avg-pool needs K=3");
  const int Ho = (H + 2 - K) / 2 + 1, Wo = (W + 2 - K) / 2 + 1;
  const int threads = threads_for(C);
  const int py = threads / (C / 8);
  const int grid = grid_for((int64_t)N * Ho * Wo, py * 4, 148 * 8);
  const size_t smem = stats ? 2 * C * sizeof(float) : 0;
  auto xx = static_cast<const __nv_bfloat16*>(x);
  auto zz = static_cast<__nv_bfloat16*>(z);
  auto pp = static_cast<__nv_bfloat16*>(pool);
  if (K == 3)
    dwconv_s2_fwd_kernel<3><<<grid, threads, smem, stream>>>(xx, x_ld, N, H, W, C, w, bias, zz, z_ld, pp, pool_ld, Ho, Wo, act, slope, stats);
  else
    dwconv_s2_fwd_kernel<4><<<grid, threads, smem, stream>>>(xx, x_ld, N, H, W, C, w, bias, zz, z_ld, pp, pool_ld, Ho, Wo, act, slope, stats);
  return check_launch("dwconv_s2_fwd");
}

int b200_dwconv_s2_dgrad(const void* dz, int dz_ld, const void* dpool, int dpool_ld, int N, int H,
                         int W, int C, int K, const float* w, void* dx, int dx_ld,
                         cudaStream_t stream) {
  if (C % 8) return set_error(B200_EINVAL, "dwconv dgrad: C=%d must be a multiple of 8", C);
  if (K != 3 && K != 4) return set_error(B200_EINVAL, "dwconv dgrad: K=%d unsupported", K);
  const int Ho = (H + 2 - K) / 2 + 1, Wo = (W + 2 - K) / 2 + 1;
  const int grid = grid_for((int64_t)N * H * W * (C / 8), 256, 148 * 16);
  auto dd = static_cast<const __nv_bfloat16*>(dz);
  auto dp = static_cast<const __nv_bfloat16*>(dpool);
  auto xx = static_cast<__nv_bfloat16*>(dx);
  if (K == 3)
    dwconv_s2_dgrad_kernel<3><<<grid, 256, 0, stream>>>(dd, dz_ld, dp, dpool_ld, N, H, W, C, Ho, Wo, w, xx, dx_ld);
  else
    dwconv_s2_dgrad_kernel<4><<<grid, 256, 0, stream>>>(dd, dz_ld, dp, dpool_ld, N, H, W, C, Ho, Wo, w, xx, dx_ld);
  return check_launch("dwconv_s2_dgrad");
}

int b200_dwconv_s2_wgrad(const void* dz, int dz_ld, const void* x, int x_ld, int N, int H, int W,
                         int C, int K, float* dw, float* dbias, cudaStream_t stream) {
  if (C % 8 || C > 1024) return set_error(B200_EINVAL, "dwconv wgrad: C=%d must be a multiple of 8 (<= 1024)", C);
  if (K != 3 && K != 4) return set_error(B200_EINVAL, "dwconv wgrad: K=%d unsupported", K);
  const int Ho = (H + 2 - K) / 2 + 1, Wo = (W + 2 - K) / 2 + 1;
  const int threads = threads_for(C);
  const int py = threads / (C / 8);
  const int grid = grid_for((int64_t)N * Ho * Wo, py * 16, 148 * 2);
  const size_t smem = (size_t)C * (K * K + 1) * sizeof(float);
  auto dd = static_cast<const __nv_bfloat16*>(dz);
  auto xx = static_cast<const __nv_bfloat16*>(x);
  if (K == 3)
    dwconv_s2_wgrad_kernel<3><<<grid, threads, smem, stream>>>(dd, dz_ld, xx, x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  else {
    static int optin = 0;
    if (!optin) {
      cudaFuncSetAttribute(dwconv_s2_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      optin = 1;
    }
    dwconv_s2_wgrad_kernel<4><<<grid, threads, smem, stream>>>(dd, dz_ld, xx, x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  }
  return check_launch("dwconv_s2_wgrad");
}

}  // extern "C"
