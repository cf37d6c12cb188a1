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

__device__ __forceinline__ void up8(const uint4& u, float (&v)[8]) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
// 16-byte load that yields zeros when `ok` is false (padding / ragged edges) without a branch
// around the load, so that all the loads of a thread can be in flight together.
__device__ __forceinline__ uint4 ldz(const __nv_bfloat16* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const uint4*>(p)) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void fma8(float (&acc)[8], const float (&v)[8], const float* wrow) {
  const float4 wa = *reinterpret_cast<const float4*>(wrow);
  const float4 wb = *reinterpret_cast<const float4*>(wrow + 4);
  acc[0] += v[0] * wa.x; acc[1] += v[1] * wa.y; acc[2] += v[2] * wa.z; acc[3] += v[3] * wa.w;
  acc[4] += v[4] * wb.x; acc[5] += v[5] * wb.y; acc[6] += v[6] * wb.z; acc[7] += v[7] * wb.w;
}

// All three kernels keep the filter transposed in shared memory ([tap][C] fp32, float4 reads) so
// that a thread's registers hold only its accumulators, and they are written for memory-level
// parallelism: every 16-byte load a thread needs for one work item is issued before the first
// dependent FMA (the first versions had the loads inside the bounds-check branches and ran at
// 0.17-0.28 of the HBM roofline inside the step's graph).

// Forward.  w: fp32 [C][K*K] (PyTorch [C,1,K,K]).  Writes z = dwconv(x) (+bias, +act) and, when
// `pool` is given, pool = avg_pool2d(x, 3, 2, 1) (only with K == 3).  stats (optional): [2][C]
// sum / sum-of-squares of the bf16-rounded z.
template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                     const float* __restrict__ w, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ z, int z_ld, __nv_bfloat16* __restrict__ pool,
                     int pool_ld, int Ho, int Wo, int act, float slope, float* __restrict__ stats) {
  extern __shared__ float s_mem[];  // [K*K][C] filter, then [2][C] statistics
  float* s_w = s_mem;
  float* s_acc = s_mem + K * K * C;
  const int groups = C >> 3;
  const int py = blockDim.x / groups;
  const int g = threadIdx.x % groups, ty = threadIdx.x / groups;
#pragma unroll 4
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) {
    const int c = i / (K * K), t = i - c * (K * K);
    s_w[t * C + c] = __ldg(w + i);
  }
  if (stats != nullptr)
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = bias != nullptr ? bias[g * 8 + j] : 0.f;
  float s1[8] = {0}, s2[8] = {0};
  const int npix = N * Ho * Wo;
  if (ty < py) {
    for (int p = blockIdx.x * py + ty; p < npix; p += gridDim.x * py) {
      const int wo = p % Wo;
      const int t2 = p / Wo;
      const int ho = t2 % Ho;
      const int n = t2 / Ho;
      const __nv_bfloat16* img = x + (size_t)n * H * W * x_ld + g * 8;
      uint4 raw[K * K];
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const int h = ho * 2 + r - 1;
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const int ww = wo * 2 + s - 1;
          const bool ok = h >= 0 && h < H && ww >= 0 && ww < W;
          raw[r * K + s] = ldz(img + ((size_t)(ok ? h : 0) * W + (ok ? ww : 0)) * x_ld, ok);
        }
      }
      float acc[8], pl[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = bs[j];
        pl[j] = 0.f;
      }
#pragma unroll
      for (int t = 0; t < K * K; ++t) {
        float v[8];
        up8(raw[t], v);
        fma8(acc, v, s_w + t * C + g * 8);
        if (pool != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pl[j] += v[j];
        }
      }
      if (act == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = acc[j] > 0.f ? acc[j] : acc[j] * slope;
      }
      st8(z + (size_t)p * z_ld + g * 8, acc);
      if (pool != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pl[j] *= (1.f / 9.f);
        st8(pool + (size_t)p * pool_ld + g * 8, pl);
      }
      if (stats != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float rr = __bfloat162float(__float2bfloat16(acc[j]));
          s1[j] += rr;
          s2[j] += rr * rr;
        }
      }
    }
  }
  if (stats != nullptr) {
    // lanes that hold the same channel group (groups < 32) are summed by shuffles first, so the
    // shared-memory atomics see one lane per group and warp
    const bool fold = groups < 32 && (32 % groups) == 0 && (blockDim.x % 32) == 0;
    if (fold) {
      for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
        }
      }
    }
    if (ty < py && (!fold || (threadIdx.x & 31) < groups)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_acc[g * 8 + j], s1[j]);
        atomicAdd(&s_acc[C + g * 8 + j], s2[j]);
      }
    }
    __syncthreads();
    flush_add_v4(stats, s_acc, 2 * C, threadIdx.x, blockDim.x);
  }
}

// ---------------------------------------------------------------- 4x4 / stride 2, shared-memory tiles
// The kernels above read every input pixel of a 4x4 stride-2 window from global memory once per output pixel that
// uses it (4x): 16 loads, 16 address computations and up to 128 L1 tag look-ups per 8 output values -- 0.17-0.34
// of HBM inside a graph on the depthwise-separable discriminators' layers (scripts/bench_small.py).  Here a CTA
// stages the (2 TH + 2) x (2 TW + 2) input pixels x CT channels under a TH x TW output tile in shared memory with
// 16-byte cp.async (zero-filled outside the map = the conv's padding), double buffered over a persistent tile
// loop, and every window element is a conflict-free LDS.128; a thread owns TWO output pixels of one channel group,
// so a filter tap's weights are read once for both.  Used for the two widest layers (C <= 64), see the launcher.
//   CT = 64: a pixel is 128 B; lanes = 8 groups of one pixel first, then the tile row.
//   CT = 32: a pixel is 64 B; lanes = 4 groups, then two tile ROWS (row pitch padded by 32 B: two input rows apart
//            = 64 B apart in bank space), then the columns.
constexpr int kDwTH = 4;                      // output rows per tile
constexpr int kDwIH = 2 * kDwTH + 2;          // input rows under it
template <int CT>
struct DwTile {
  static constexpr int G = CT / 8;                        // 16-byte channel groups per pixel
  static constexpr int TW = 512 / (kDwTH * G);            // output columns per tile (512 (pixel, group) items)
  static constexpr int IW = 2 * TW + 2;                   // input columns under it
  static constexpr int PITCH = IW * CT * 2 + (CT == 32 ? 32 : 0);   // bytes per staged input row
  static constexpr int BYTES = kDwIH * PITCH;
  static constexpr int PIECES = kDwIH * IW * G;           // 16-byte pieces per tile
};

template <int CT>
__device__ __forceinline__ void dw4_stage(uint32_t dst, const __nv_bfloat16* __restrict__ x, int x_ld, int n, int H, int W,
                                          int c0, int h_in0, int w_in0) {
  using T = DwTile<CT>;
  for (int i = threadIdx.x; i < T::PIECES; i += blockDim.x) {
    const int g = i % T::G;
    const int col = (i / T::G) % T::IW;
    const int row = i / (T::G * T::IW);
    const int h = h_in0 + row, w = w_in0 + col;
    const bool ok = h >= 0 && h < H && w >= 0 && w < W;
    const __nv_bfloat16* src = x + (((size_t)n * H + (ok ? h : 0)) * W + (ok ? w : 0)) * x_ld + c0 + g * 8;
    cp_async16(dst + row * T::PITCH + col * (CT * 2) + g * 16, src, ok ? 16u : 0u);
  }
}

// z = act(dwconv4x4s2p1(x) + bias) (+ BatchNorm statistics of the stored values).  C % CT == 0.
template <int CT>
__global__ void __launch_bounds__(256, 2)
dw4_fwd_tiled_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                     const float* __restrict__ w, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ z, int z_ld, int Ho, int Wo, int act, float slope,
                     float* __restrict__ stats) {
  using T = DwTile<CT>;
  extern __shared__ __align__(16) uint8_t s_raw[];
  uint8_t* s_tile = s_raw;                                          // two input tiles
  float* s_w = reinterpret_cast<float*>(s_raw + 2 * T::BYTES);      // [16][CT] filter of the current channel chunk
  float* s_b = s_w + 16 * CT;                                       // [CT] bias
  float* s_acc = s_b + CT;                                          // [2][CT] statistics of the current chunk
  const int tiles_w = (Wo + T::TW - 1) / T::TW, tiles_h = (Ho + kDwTH - 1) / kDwTH;
  const int chunks = C / CT;
  const int per_chunk = N * tiles_h * tiles_w;
  const int total = chunks * per_chunk;
  // item = (group, pixel of the tile); the thread's two items are 256 apart: same group (256 % G == 0)
  int g, ty[2], tx[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int it = threadIdx.x + 256 * j;
    if (CT == 64) {
      g = it % T::G;
      tx[j] = (it / T::G) % T::TW;
      ty[j] = it / (T::G * T::TW);
    } else {
      g = it % T::G;
      const int rp = (it / T::G) % 2;                 // row of a row pair
      tx[j] = (it / (T::G * 2)) % T::TW;
      ty[j] = 2 * (it / (T::G * 2 * T::TW)) + rp;
    }
  }
  auto coords = [&](int t, int* cc, int* n, int* h0, int* w0) {
    *cc = t / per_chunk;
    int r = t - *cc * per_chunk;
    *n = r / (tiles_h * tiles_w);
    r -= *n * tiles_h * tiles_w;
    *h0 = (r / tiles_w) * kDwTH;
    *w0 = (r % tiles_w) * T::TW;
  };
  const uint32_t tile_u32 = smem_u32(s_tile);
  float s1[8] = {0}, s2[8] = {0};
  int cur_cc = -1;
  auto flush_stats = [&]() {   // this thread's partial sums -> the chunk's shared sums -> global
    for (int off = T::G; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
      }
    }
    if ((threadIdx.x & 31) < T::G) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_acc[g * 8 + j], s1[j]);
        atomicAdd(&s_acc[CT + g * 8 + j], s2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < CT; i += blockDim.x) {
      atomicAdd(&stats[cur_cc * CT + i], s_acc[i]);
      atomicAdd(&stats[C + cur_cc * CT + i], s_acc[CT + i]);
    }
  };
  int t = blockIdx.x, buf = 0;
  if (t < total) {
    int cc, n, h0, w0;
    coords(t, &cc, &n, &h0, &w0);
    dw4_stage<CT>(tile_u32, x, x_ld, n, H, W, cc * CT, 2 * h0 - 1, 2 * w0 - 1);
  }
  cp_async_commit();
  for (; t < total; t += gridDim.x, buf ^= 1) {
    int cc, n, h0, w0;
    coords(t, &cc, &n, &h0, &w0);
    const int tn = t + gridDim.x;
    if (tn < total) {
      int cc2, n2, h2, w2;
      coords(tn, &cc2, &n2, &h2, &w2);
      dw4_stage<CT>(tile_u32 + (buf ^ 1) * T::BYTES, x, x_ld, n2, H, W, cc2 * CT, 2 * h2 - 1, 2 * w2 - 1);
    }
    cp_async_commit();
    if (cc != cur_cc) {   // (block-uniform) next channel chunk: its filter, bias, statistics
      if (stats != nullptr && cur_cc >= 0) flush_stats();
      __syncthreads();
      for (int i = threadIdx.x; i < 16 * CT; i += blockDim.x) {
        const int c = i % CT, tap = i / CT;
        s_w[i] = __ldg(w + (size_t)(cc * CT + c) * 16 + tap);
      }
      for (int i = threadIdx.x; i < CT; i += blockDim.x) {
        s_b[i] = bias != nullptr ? __ldg(bias + cc * CT + i) : 0.f;
        s_acc[i] = 0.f;
        s_acc[CT + i] = 0.f;
      }
      cur_cc = cc;
    }
    cp_async_wait<1>();    // this tile has landed (the next one may still be in flight)
    __syncthreads();
    const uint8_t* tile = s_tile + buf * T::BYTES;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] = s_b[g * 8 + c];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float* wrow = s_w + (r * 4 + s) * CT + g * 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint4 raw = *reinterpret_cast<const uint4*>(tile + (2 * ty[j] + r) * T::PITCH + (2 * tx[j] + s) * (CT * 2) + g * 16);
          float v[8];
          up8(raw, v);
          fma8(acc[j], v, wrow);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ho = h0 + ty[j], wo = w0 + tx[j];
      if (ho < Ho && wo < Wo) {
        if (act == 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[j][c] = fmaxf(acc[j][c], 0.f) + slope * fminf(acc[j][c], 0.f);
        } else if (act == 1) {
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[j][c] = fmaxf(acc[j][c], 0.f);
        }
        st8(z + (((size_t)n * Ho + ho) * Wo + wo) * z_ld + cc * CT + g * 8, acc[j]);
        if (stats != nullptr) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float rr = __bfloat162float(__float2bfloat16(acc[j][c]));
            s1[c] += rr;
            s2[c] += rr * rr;
          }
        }
      }
    }
    __syncthreads();   // everyone is done with this buffer before the next iteration's loads overwrite it
  }
  cp_async_wait<0>();
  if (stats != nullptr && cur_cc >= 0) flush_stats();
}

// Data gradient: dx[n,h,w,c] = sum_{r,s} dz[n,(h+1-r)/2,(w+1-s)/2,c] * w[c,r,s]
//                              (+ sum over the 3x3/s2 pooling windows covering (h,w) of dpool / 9)
// One work item = the 2x2 block of dx pixels (2i+ph, 2j+pw): which filter tap links an output
// pixel to dz[i+di, j+dj] depends only on the parities (r = ph + 1 - 2 di, s = pw + 1 - 2 dj), so the
// control flow is uniform, the (2 or 3)^2 dz pixels the block needs are loaded up front and every
// tap test below is resolved at compile time.
template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_s2_dgrad_kernel(const __nv_bfloat16* __restrict__ dz, int dz_ld,
                       const __nv_bfloat16* __restrict__ dpool, int dpool_ld, int N, int H, int W,
                       int C, int Ho, int Wo, const float* __restrict__ w,
                       __nv_bfloat16* __restrict__ dx, int dx_ld) {
  extern __shared__ float s_w[];  // [K*K][C]
  constexpr int DLO = (K == 4) ? -1 : 0;   // dz offsets di, dj in [DLO, 1]
  constexpr int ND = 2 - DLO;
  const int groups = C >> 3;
  const int py = blockDim.x / groups;
  const int g = threadIdx.x % groups, ty = threadIdx.x / groups;
#pragma unroll 4
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) {
    const int c = i / (K * K), t = i - c * (K * K);
    s_w[t * C + c] = __ldg(w + i);
  }
  __syncthreads();
  if (ty >= py) return;
  const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;
  const int nblk = N * Hb * Wb;
  for (int p = blockIdx.x * py + ty; p < nblk; p += gridDim.x * py) {
    const int j = p % Wb;
    const int t2 = p / Wb;
    const int i = t2 % Hb;
    const int n = t2 / Hb;
    uint4 rz[ND][ND], rp[ND][ND];
#pragma unroll
    for (int a = 0; a < ND; ++a) {
#pragma unroll
      for (int b = 0; b < ND; ++b) {
        const int ho = i + DLO + a, wo = j + DLO + b;
        const bool ok = ho >= 0 && ho < Ho && wo >= 0 && wo < Wo;
        const size_t q = ((size_t)n * Ho + (ok ? ho : 0)) * Wo + (ok ? wo : 0);
        rz[a][b] = ldz(dz + q * dz_ld + g * 8, ok);
        if (dpool != nullptr) rp[a][b] = ldz(dpool + q * dpool_ld + g * 8, ok);
      }
    }
    float acc[2][2][8];
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[ph][pw][c] = 0.f;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
#pragma unroll
      for (int b = 0; b < ND; ++b) {
        float v[8], u[8];
        up8(rz[a][b], v);
        if (dpool != nullptr) up8(rp[a][b], u);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const int r = ph + 1 - 2 * (DLO + a);
          if (r < 0 || r >= K) continue;
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const int s = pw + 1 - 2 * (DLO + b);
            if (s < 0 || s >= K) continue;
            fma8(acc[ph][pw], v, s_w + (r * K + s) * C + g * 8);
            if (dpool != nullptr) {   // the pooling window is the 3x3 tap set with weight 1/9 (K == 3 only)
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[ph][pw][c] += u[c] * (1.f / 9.f);
            }
          }
        }
      }
    }
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        const int h = 2 * i + ph, ww = 2 * j + pw;
        if (h < H && ww < W) st8(dx + (((size_t)n * H + h) * W + ww) * dx_ld + g * 8, acc[ph][pw]);
      }
    }
  }
}

// Filter / bias gradient: dw[c][r][s] += sum_pixels dz[.,c] * x[.*2 + tap, c]; dbias[c] += sum dz.
// A thread owns (channel group, filter row r): K accumulators x 8 channels; the K threads of a
// pixel share the dz load through L1.  Two pixels per iteration, all 2 (K + 1) loads issued first.
template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_s2_wgrad_kernel(const __nv_bfloat16* __restrict__ dz, int dz_ld,
                       const __nv_bfloat16* __restrict__ x, int x_ld, int N, int H, int W, int C,
                       int Ho, int Wo, float* __restrict__ dw, float* __restrict__ dbias) {
  extern __shared__ float s_acc[];  // [C][K*K + 1]
  const int groups = C >> 3;
  const int slots = groups * K;                 // (group, filter row)
  const int py = blockDim.x / slots;
  const int slot = threadIdx.x % slots, ty = threadIdx.x / slots;
  const int g = slot % groups, r = slot / groups;
  constexpr int KK1 = K * K + 1;
  constexpr int PP = 2;
  for (int i = threadIdx.x; i < C * KK1; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float acc[K][8];
  float ab[8] = {0};
#pragma unroll
  for (int t = 0; t < K; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int npix = N * Ho * Wo;
  if (ty < py) {
    for (int p0 = (blockIdx.x * py + ty) * PP; p0 < npix; p0 += gridDim.x * py * PP) {
      uint4 rd[PP], rx[PP][K];
#pragma unroll
      for (int u = 0; u < PP; ++u) {
        const int p = p0 + u;
        const bool live = p < npix;
        const int pc = live ? p : 0;
        const int wo = pc % Wo;
        const int t2 = pc / Wo;
        const int ho = t2 % Ho;
        const int n = t2 / Ho;
        rd[u] = ldz(dz + (size_t)pc * dz_ld + g * 8, live);
        const int h = ho * 2 + r - 1;
        const bool hok = live && h >= 0 && h < H;
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const int ww = wo * 2 + s - 1;
          const bool ok = hok && ww >= 0 && ww < W;
          rx[u][s] = ldz(x + (((size_t)n * H + (ok ? h : 0)) * W + (ok ? ww : 0)) * x_ld + g * 8, ok);
        }
      }
#pragma unroll
      for (int u = 0; u < PP; ++u) {
        float d[8];
        up8(rd[u], d);
        if (r == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ab[j] += d[j];
        }
#pragma unroll
        for (int s = 0; s < K; ++s) {
          float v[8];
          up8(rx[u][s], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[s][j] += d[j] * v[j];
        }
      }
    }
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(g * 8 + j) * KK1 + r * K + s], acc[s][j]);
    if (r == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(g * 8 + j) * KK1 + K * K], ab[j]);
    }
  }
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
  if (C % 8 || C > 1024) return set_error(B200_EINVAL, "dwconv: C=%d must be a multiple of 8 (<= 1024)", C);
  if (K != 3 && K != 4) return set_error(B200_EINVAL, "dwconv: K=%d unsupported", K);
  if (pool != nullptr && K != 3) return set_error(B200_EINVAL, "dwconv: fused avg-pool needs K=3");
  const int Ho = (H + 2 - K) / 2 + 1, Wo = (W + 2 - K) / 2 + 1;
  if ((int64_t)N * H * W >= (1ll << 31)) return set_error(B200_EINVAL, "dwconv: tensor too large");
  const int threads = threads_for(C);
  const int py = threads / (C / 8);
  const int grid = grid_for((int64_t)N * Ho * Wo, py * 2, 148 * 2);   // resident CTAs loop: few end-of-CTA reductions
  const size_t smem = (size_t)(K * K + 2) * C * sizeof(float);
  static int optin = 0;
  if (!optin) {
    cudaFuncSetAttribute(dwconv_s2_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(dwconv_s2_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(dwconv_s2_dgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(dwconv_s2_dgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(dwconv_s2_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(dwconv_s2_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    optin = 1;
  }
  auto xx = static_cast<const __nv_bfloat16*>(x);
  auto zz = static_cast<__nv_bfloat16*>(z);
  auto pp = static_cast<__nv_bfloat16*>(pool);
  if (K == 4 && pool == nullptr && (C == 32 || C == 64) && x_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // shared-memory tiled kernel (two CTAs of ~92 KB per SM, persistent over the tiles) on the wide, thin layers:
    // 19(32) channels at 512x1024 148.7 -> 121.1 us, 64 at 258x514 87.6 -> 79.6 us inside a graph; from 128
    // channels up (maps of <= 131x259) the per-pixel kernel is as fast or faster (47.9 / 28.1 / 21.1 us against
    // 46.0 / 31.6 / 22.6).  What bounds the tiled kernel is shared-memory bandwidth: 256 B of window data plus
    // 256 B of filter taps per (pixel, group) item -- four items per thread and a sliding window would halve it.
    static int optin4 = 0;
    if (!optin4) {
      cudaFuncSetAttribute(dw4_fwd_tiled_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(dw4_fwd_tiled_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      optin4 = 1;
    }
    const int ct = C == 32 ? 32 : 64;
    const int tw = ct == 32 ? DwTile<32>::TW : DwTile<64>::TW;
    const long tiles = (long)(C / ct) * N * ((Ho + kDwTH - 1) / kDwTH) * ((Wo + tw - 1) / tw);
    const int tgrid = (int)(tiles < 296 ? tiles : 296);
    const size_t tsmem = 2 * (size_t)(ct == 32 ? DwTile<32>::BYTES : DwTile<64>::BYTES) + (size_t)(16 + 1 + 2) * ct * sizeof(float);
    if (ct == 32)
      dw4_fwd_tiled_kernel<32><<<tgrid, 256, tsmem, stream>>>(xx, x_ld, N, H, W, C, w, bias, zz, z_ld, Ho, Wo, act, slope, stats);
    else
      dw4_fwd_tiled_kernel<64><<<tgrid, 256, tsmem, stream>>>(xx, x_ld, N, H, W, C, w, bias, zz, z_ld, Ho, Wo, act, slope, stats);
    return check_launch("dwconv_s2_fwd(tiled)");
  }
  if (K == 3)
    dwconv_s2_fwd_kernel<3><<<grid, threads, smem, stream>>>(xx, x_ld, N, H, W, C, w, bias, zz, z_ld, pp, pool_ld, Ho, Wo, act, slope, stats);
  else
    dwconv_s2_fwd_kernel<4><<<grid, threads, smem, stream>>>(xx, x_ld, N, H, W, C, w, bias, zz, z_ld, pp, pool_ld, Ho, Wo, act, slope, stats);
  return check_launch("dwconv_s2_fwd");
}

int b200_dwconv_s2_dgrad(const void* dz, int dz_ld, const void* dpool, int dpool_ld, int N, int H,
                         int W, int C, int K, const float* w, void* dx, int dx_ld,
                         cudaStream_t stream) {
  if (C % 8 || C > 1024) return set_error(B200_EINVAL, "dwconv dgrad: C=%d must be a multiple of 8 (<= 1024)", C);
  if (K != 3 && K != 4) return set_error(B200_EINVAL, "dwconv dgrad: K=%d unsupported", K);
  if ((int64_t)N * H * W >= (1ll << 31)) return set_error(B200_EINVAL, "dwconv dgrad: tensor too large");
  const int Ho = (H + 2 - K) / 2 + 1, Wo = (W + 2 - K) / 2 + 1;
  const int threads = threads_for(C);
  const int py = threads / (C / 8);
  const int grid = grid_for((int64_t)N * ((H + 1) / 2) * ((W + 1) / 2), py * 2, 148 * 16);
  const size_t smem = (size_t)K * K * C * sizeof(float);
  if (smem > 48 * 1024) {
    static int optin = 0;
    if (!optin) {
      cudaFuncSetAttribute(dwconv_s2_dgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      cudaFuncSetAttribute(dwconv_s2_dgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      optin = 1;
    }
  }
  auto dd = static_cast<const __nv_bfloat16*>(dz);
  auto dp = static_cast<const __nv_bfloat16*>(dpool);
  auto xx = static_cast<__nv_bfloat16*>(dx);
  if (K == 3)
    dwconv_s2_dgrad_kernel<3><<<grid, threads, smem, stream>>>(dd, dz_ld, dp, dpool_ld, N, H, W, C, Ho, Wo, w, xx, dx_ld);
  else
    dwconv_s2_dgrad_kernel<4><<<grid, threads, smem, stream>>>(dd, dz_ld, dp, dpool_ld, N, H, W, C, Ho, Wo, w, xx, dx_ld);
  return check_launch("dwconv_s2_dgrad");
}

int b200_dwconv_s2_wgrad(const void* dz, int dz_ld, const void* x, int x_ld, int N, int H, int W,
                         int C, int K, float* dw, float* dbias, cudaStream_t stream) {
  if (C % 8 || C > 1024) return set_error(B200_EINVAL, "dwconv wgrad: C=%d must be a multiple of 8 (<= 1024)", C);
  if (K != 3 && K != 4) return set_error(B200_EINVAL, "dwconv wgrad: K=%d unsupported", K);
  if ((int64_t)N * H * W >= (1ll << 31)) return set_error(B200_EINVAL, "dwconv wgrad: tensor too large");
  const int Ho = (H + 2 - K) / 2 + 1, Wo = (W + 2 - K) / 2 + 1;
  const int slots = (C / 8) * K;
  int threads = (256 / slots) * slots;
  if (threads == 0) threads = slots;   // C = 1024, K = 4: 512 threads
  if (threads > 1024) return set_error(B200_EINVAL, "dwconv wgrad: C=%d too wide", C);
  const int py = threads / slots;
  const int grid = grid_for((int64_t)N * Ho * Wo, py * 16, 148 * 2);
  const size_t smem = (size_t)C * (K * K + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    static int optin = 0;
    if (!optin) {
      cudaFuncSetAttribute(dwconv_s2_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      cudaFuncSetAttribute(dwconv_s2_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      optin = 1;
    }
  }
  auto dd = static_cast<const __nv_bfloat16*>(dz);
  auto xx = static_cast<const __nv_bfloat16*>(x);
  if (K == 3)
    dwconv_s2_wgrad_kernel<3><<<grid, threads, smem, stream>>>(dd, dz_ld, xx, x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  else
    dwconv_s2_wgrad_kernel<4><<<grid, threads, smem, stream>>>(dd, dz_ld, xx, x_ld, N, H, W, C, Ho, Wo, dw, dbias);
  return check_launch("dwconv_s2_wgrad");
}

}  // extern "C"
