// BatchNorm / activation passes over bf16 NHWC activations (HBM-bound, 16-byte vector accesses).
//   reference: nn.BatchNorm2d + ReLU in ConvX (model/stdcnet.py:9-14) and ConvBNReLU
//   (model/model_stages.py:21-29); BatchNorm2d + LeakyReLU(0.2) in DepthWiseSepBNFCDiscriminator
//   (model/discriminator.py:103-131); LeakyReLU(0.2) after biased convs (discriminator.py:17-26).
//
// A "view" is (base pointer, pixel stride ld in elements): channels [0, C) of each of `npix`
// pixels.  All C are multiples of 8.
#include <cooperative_groups.h>
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 ldraw(const __nv_bfloat16* p) {
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                            pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : v * slope;
  return v;
}
__device__ __forceinline__ float act_grad(float pre, int act, float slope) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) return pre > 0.f ? 1.f : slope;
  return 1.f;
}

// All streaming kernels below share one shape: a block is (C/8 channel groups) x PY pixel rows,
// a thread keeps its channel group for the whole kernel (per-channel parameters live in registers)
// and walks pixels in batches of UNR independent 16-byte loads per tensor (memory-level parallelism).
constexpr int UNR = 4;   // single-tensor kernels
constexpr int UNR3 = 2;  // kernels streaming three tensors
constexpr int UNRF = 2;  // the fused BatchNorm backward: two batches of UNRF x 3 loads live per thread (double buffering) ...
constexpr int UNRF1 = 4; // ... or of UNRF1 x 2 loads when there is one gradient input
constexpr int kRedRep = 8;  // replicas of the partial-sum buffer of the fused BatchNorm backward

struct Slot {
  int g, ty, py;
};
__device__ __forceinline__ Slot slot_of(int C) {
  Slot s;
  const int groups = C >> 3;
  s.py = blockDim.x / groups;
  s.g = threadIdx.x % groups;
  s.ty = threadIdx.x / groups;
  return s;
}

// ---------------------------------------------------------------- statistics
// stats[0][c] += sum x, stats[1][c] += sum x^2 over all pixels.
__global__ void __launch_bounds__(256)
channel_stats_kernel(const __nv_bfloat16* __restrict__ x, int ld, int C, int64_t npix,
                     float* __restrict__ stats) {
  extern __shared__ float s_acc[];  // [2][C]
  const Slot t = slot_of(C);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float s1[8] = {0}, s2[8] = {0};
  const int64_t step = (int64_t)gridDim.x * t.py * UNR;
  for (int64_t base = (int64_t)blockIdx.x * t.py * UNR; base < npix; base += step) {
    float v[UNR][8];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      if (p < npix) load8(x + p * ld + t.g * 8, v[u]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += v[u][j];
        s2[j] += v[u][j] * v[u][j];
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_acc[t.g * 8 + j], s1[j]);
    atomicAdd(&s_acc[C + t.g * 8 + j], s2[j]);
  }
  __syncthreads();
  flush_add_v4(stats, s_acc, 2 * C, threadIdx.x, blockDim.x);
}

// mean / biased var from (sum, sumsq) -> scale = gamma*rstd, shift = beta - mean*scale; running
// statistics updated with momentum (unbiased variance), exactly nn.BatchNorm2d's train step.
// training == 0: scale/shift from the running statistics (eval mode).
__device__ __forceinline__ void bn_coeffs(const float* stats, int C, int c, float count,
                                          const float* gamma, const float* beta,
                                          const float* running_mean, const float* running_var,
                                          float eps, int training, float* mean, float* var,
                                          float* rstd, float* scale, float* shift) {
  if (training) {
    *mean = stats[c] / count;
    *var = fmaxf(stats[C + c] / count - *mean * *mean, 0.f);
  } else {
    *mean = running_mean[c];
    *var = running_var[c];
  }
  *rstd = rsqrtf(*var + eps);
  const float g = gamma != nullptr ? gamma[c] : 1.f;
  const float b = beta != nullptr ? beta[c] : 0.f;
  *scale = g * *rstd;
  *shift = b - *mean * g * *rstd;
}

__global__ void bn_finalize_kernel(const float* __restrict__ stats, int C, float count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, int training,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var, rstd, sc, sh;
  bn_coeffs(stats, C, c, count, gamma, beta, running_mean, running_var, eps, training, &mean, &var,
            &rstd, &sc, &sh);
  if (training && running_mean != nullptr) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
  scale[c] = sc;
  shift[c] = sh;
  if (mean_out != nullptr) mean_out[c] = mean;
  if (rstd_out != nullptr) rstd_out[c] = rstd;
}

// ------------------------------------------------------------- normalise + act
// y = act(x * scale[c] + shift[c]).  Two ways to get scale/shift:
//   stats == nullptr : read them from `scale` / `shift` (nullptr: identity)
//   stats != nullptr : finalise BatchNorm in-kernel from (sum, sumsq) — every thread derives the
//                      coefficients of its own 8 channels; block 0 also publishes scale / shift /
//                      mean / rstd for the backward pass and updates the running statistics.
__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y,
                    int y_ld, int C, int64_t npix, float* __restrict__ scale,
                    float* __restrict__ shift, int act, float slope,
                    const float* __restrict__ stats, float count, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ running_mean,
                    float* __restrict__ running_var, float momentum, float eps, int training,
                    float* __restrict__ mean_out, float* __restrict__ rstd_out, int finalize) {
  const Slot t = slot_of(C);
  const int64_t first = (int64_t)blockIdx.x * t.py * UNR;
  uint4 cur[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int64_t p = first + u * t.py + t.ty;
    cur[u] = p < npix ? ldraw(x + p * x_ld + t.g * 8) : make_uint4(0u, 0u, 0u, 0u);
  }
  float sc[8], sh[8];
  if (finalize) {
    // per-channel vectors of this thread's 8 channels with 16-byte loads (the coefficient set-up is a
    // latency chain every thread runs before its first pixel: it is most of a small layer's time)
    float a_[8], b_[8], g_[8], be_[8];
    const int c0 = t.g * 8;
    auto ld8 = [&](const float* p, float (&o)[8], float dflt) {
      if (p != nullptr) {
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(p + c0)), u1 = __ldg(reinterpret_cast<const float4*>(p + c0) + 1);
        o[0] = u0.x; o[1] = u0.y; o[2] = u0.z; o[3] = u0.w; o[4] = u1.x; o[5] = u1.y; o[6] = u1.z; o[7] = u1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = dflt;
      }
    };
    if (training) {
      ld8(stats, a_, 0.f);
      ld8(stats + C, b_, 0.f);
    } else {
      ld8(running_mean, a_, 0.f);
      ld8(running_var, b_, 1.f);
    }
    ld8(gamma, g_, 1.f);
    ld8(beta, be_, 0.f);
    const bool publish = blockIdx.x == 0 && t.ty == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float mean, var;
      if (training) {
        mean = a_[j] / count;
        var = fmaxf(b_[j] / count - mean * mean, 0.f);
      } else {
        mean = a_[j];
        var = b_[j];
      }
      const float rstd = rsqrtf(var + eps);
      sc[j] = g_[j] * rstd;
      sh[j] = be_[j] - mean * g_[j] * rstd;
      if (publish) {
        if (scale != nullptr) {
          scale[c] = sc[j];
          shift[c] = sh[j];
        }
        if (mean_out != nullptr) {
          mean_out[c] = mean;
          rstd_out[c] = rstd;
        }
        if (training && running_mean != nullptr) {
          const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = t.g * 8 + j;
      sc[j] = scale != nullptr ? scale[c] : 1.f;
      sh[j] = shift != nullptr ? shift[c] : 0.f;
    }
  }
  const int64_t step = (int64_t)gridDim.x * t.py * UNR;
  for (int64_t base = first; base < npix; base += step) {
    // the next batch's loads are issued before this batch is touched (and the first batch's before the
    // coefficient chain above: on a small layer that chain and the first loads ARE the kernel)
    uint4 nxt[UNR];
    const int64_t nb = base + step;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t p = nb + u * t.py + t.ty;
      nxt[u] = p < npix ? ldraw(x + p * x_ld + t.g * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      if (p < npix) {
        float v[8];
        unpack8(cur[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_fwd(v[j] * sc[j] + sh[j], act, slope);
        store8(y + p * y_ld + t.g * 8, v);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) cur[u] = nxt[u];
  }
}

// ------------------------------------------------------------------ backward
// g = (dy1 + dy2) * act'(z*scale + shift);  xhat = (z - mean) * rstd
// red[0][c] += sum g (= dbeta), red[1][c] += sum g*xhat (= dgamma)
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy1, int dy1_ld,
                         const __nv_bfloat16* __restrict__ dy2, int dy2_ld,
                         const __nv_bfloat16* __restrict__ z, int z_ld, int C, int64_t npix,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ rstd, int act,
                         float slope, float* __restrict__ red) {
  extern __shared__ float s_acc[];  // [2][C]
  const Slot t = slot_of(C);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float sc[8], sh[8], a1[8] = {0}, a2[8] = {0};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[t.g * 8 + j];
    sh[j] = shift[t.g * 8 + j];
  }
  const int64_t step = (int64_t)gridDim.x * t.py * UNR3;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int64_t base = (int64_t)blockIdx.x * t.py * UNR3; base < npix; base += step) {
    uint4 rd[UNR3], rz[UNR3], re[UNR3];
#pragma unroll
    for (int u = 0; u < UNR3; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      const bool ok = p < npix;
      rd[u] = ok ? ldraw(dy1 + p * dy1_ld + t.g * 8) : zero4;  // zero gradient: contributes nothing
      rz[u] = ok ? ldraw(z + p * z_ld + t.g * 8) : zero4;
      re[u] = (ok && dy2 != nullptr) ? ldraw(dy2 + p * dy2_ld + t.g * 8) : zero4;
    }
#pragma unroll
    for (int u = 0; u < UNR3; ++u) {
      float d[8], zz[8], e[8];
      unpack8(rd[u], d);
      unpack8(rz[u], zz);
      unpack8(re[u], e);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = (d[j] + e[j]) * act_grad(zz[j] * sc[j] + sh[j], act, slope);
        a1[j] += gg;
        a2[j] += gg * zz[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)  // sum g*xhat = rstd * (sum g*z - mean * sum g)
    a2[j] = (a2[j] - mean[t.g * 8 + j] * a1[j]) * rstd[t.g * 8 + j];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_acc[t.g * 8 + j], a1[j]);
    atomicAdd(&s_acc[C + t.g * 8 + j], a2[j]);
  }
  __syncthreads();
  flush_add_v4(red, s_acc, 2 * C, threadIdx.x, blockDim.x);
}

// dz = scale * (g - red0/M - xhat * red1/M)        (train-mode BatchNorm backward)
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy1, int dy1_ld,
                        const __nv_bfloat16* __restrict__ dy2, int dy2_ld,
                        const __nv_bfloat16* __restrict__ z, int z_ld,
                        __nv_bfloat16* __restrict__ dz, int dz_ld, int C, int64_t npix,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ red, float inv_count, int act, float slope) {
  const Slot t = slot_of(C);
  // dz = sc*g - sc*r0 - sc*r1*rs*(z - mu) = sc*g + k1*z + k0
  float sc[8], sh[8], k0[8], k1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = t.g * 8 + j;
    sc[j] = scale[c];
    sh[j] = shift[c];
    const float r0 = red[c] * inv_count, r1 = red[C + c] * inv_count;
    k1[j] = -sc[j] * r1 * rstd[c];
    k0[j] = -sc[j] * r0 - k1[j] * mean[c];
  }
  const int64_t step = (int64_t)gridDim.x * t.py * UNR3;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int64_t base = (int64_t)blockIdx.x * t.py * UNR3; base < npix; base += step) {
    uint4 rd[UNR3], rz[UNR3], re[UNR3];
#pragma unroll
    for (int u = 0; u < UNR3; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      const bool ok = p < npix;
      rd[u] = ok ? ldraw(dy1 + p * dy1_ld + t.g * 8) : zero4;
      rz[u] = ok ? ldraw(z + p * z_ld + t.g * 8) : zero4;
      re[u] = (ok && dy2 != nullptr) ? ldraw(dy2 + p * dy2_ld + t.g * 8) : zero4;
    }
#pragma unroll
    for (int u = 0; u < UNR3; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      if (p < npix) {
        float d[8], zz[8], e[8], o[8];
        unpack8(rd[u], d);
        unpack8(rz[u], zz);
        unpack8(re[u], e);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = (d[j] + e[j]) * act_grad(zz[j] * sc[j] + sh[j], act, slope);
          o[j] = sc[j] * gg + k1[j] * zz[j] + k0[j];
        }
        store8(dz + p * dz_ld + t.g * 8, o);
      }
    }
  }
}

// Both backward passes in ONE cooperative launch: reductions, grid-wide barrier, then dz.  The
// second pass walks the pixels in reverse so that it starts on the lines the first pass touched
// last (for all but the largest layers dy and z are still L2-resident: no second HBM read).
//
// Both passes are software-pipelined: the 16-byte loads of batch i+1 are issued before batch i is
// reduced / written, so a thread always has loads in flight (the un-pipelined loop paid one full
// memory latency per batch: 0.30-0.46 of the HBM rate on the large layers, measured with
// scripts/bench_small.py).  After the grid barrier ONE pass over the kRedRep partial-sum replicas per
// CTA rebuilds the totals in shared memory (every thread re-reading all replicas for its 8 channels
// cost 128 L2 loads per thread -- the 12-25 us floor of the small layers).
// Bytes in flight bound these passes (512 threads per SM x loads per thread x 16 B against ~1.2 us of loaded HBM
// latency): a layer with ONE gradient input (most of them) has no use for the registers of the second one and
// runs UNRF1 = 4 pixels per batch instead of UNRF = 2 -- 64 KB per SM in flight instead of 32.
template <bool DY2, int U>
struct BwdBatch {
  uint4 d[U], z[U], e[DY2 ? U : 1];
};

template <bool DY2, int U>
__device__ __forceinline__ void bwd_load(BwdBatch<DY2, U>& b, const __nv_bfloat16* dy1, int dy1_ld, const __nv_bfloat16* dy2,
                                         int dy2_ld, const __nv_bfloat16* z, int z_ld, int64_t base, int64_t npix,
                                         const Slot& t) {
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t p = base + u * t.py + t.ty;
    const bool ok = base >= 0 && p < npix;
    b.d[u] = ok ? ldraw(dy1 + p * dy1_ld + t.g * 8) : zero4;  // zero gradient: contributes nothing
    b.z[u] = ok ? ldraw(z + p * z_ld + t.g * 8) : zero4;
    if (DY2) b.e[u] = ok ? ldraw(dy2 + p * dy2_ld + t.g * 8) : zero4;
  }
}

template <bool DY2, int U>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy1, int dy1_ld,
                        const __nv_bfloat16* __restrict__ dy2, int dy2_ld,
                        const __nv_bfloat16* __restrict__ z, int z_ld,
                        __nv_bfloat16* __restrict__ dz, int dz_ld, int C, int64_t npix,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        float* red, float inv_count, int act, float slope) {
  extern __shared__ float s_acc[];  // [2][C]
  const Slot t = slot_of(C);
  const int groups = C >> 3;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float sc[8], sh[8], a1[8] = {0}, a2[8] = {0};
  {
    const float4* s4 = reinterpret_cast<const float4*>(scale + t.g * 8);
    const float4* h4 = reinterpret_cast<const float4*>(shift + t.g * 8);
    const float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1), h0 = __ldg(h4), h1 = __ldg(h4 + 1);
    sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
    sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
  }
  const int64_t step = (int64_t)gridDim.x * t.py * U;
  const int64_t first = (int64_t)blockIdx.x * t.py * U;
  int64_t last_base = -1;
  {
    BwdBatch<DY2, U> cur, nxt;
    bwd_load(cur, dy1, dy1_ld, dy2, dy2_ld, z, z_ld, first < npix ? first : -1, npix, t);
    for (int64_t base = first; base < npix; base += step) {
      last_base = base;
      const int64_t nb = base + step;
      bwd_load(nxt, dy1, dy1_ld, dy2, dy2_ld, z, z_ld, nb < npix ? nb : -1, npix, t);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float d[8], zz[8], e[8];
        unpack8(cur.d[u], d);
        unpack8(cur.z[u], zz);
        if (DY2) {
          unpack8(cur.e[u], e);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] += e[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = d[j] * act_grad(zz[j] * sc[j] + sh[j], act, slope);
          a1[j] += gg;
          a2[j] += gg * zz[j];
        }
      }
      cur = nxt;
    }
  }
  // block reduction: lanes holding the same channel group fold by shuffles, then one shared-memory
  // atomic per (warp, channel)
  const bool fold = groups < 32 && (32 % groups) == 0 && (blockDim.x % 32) == 0;
  if (fold) {
    for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], off);
        a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], off);
      }
    }
  }
  if (!fold || (threadIdx.x & 31) < groups) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a2[j] = (a2[j] - mean[t.g * 8 + j] * a1[j]) * rstd[t.g * 8 + j];
      atomicAdd(&s_acc[t.g * 8 + j], a1[j]);
      atomicAdd(&s_acc[C + t.g * 8 + j], a2[j]);
    }
  }
  __syncthreads();
  // Same-address global atomics serialise at ~14 ns each in L2, so the CTAs spread their partial
  // sums over kRedRep replicas (red[1 + rep][2][C]); the totals are rebuilt after the grid barrier.
  float* rep = red + (size_t)(1 + blockIdx.x % kRedRep) * 2 * C;
  flush_add_v4(rep, s_acc, 2 * C, threadIdx.x, blockDim.x);
  __threadfence();
  cooperative_groups::this_grid().sync();
  // totals: one pass over the replicas per CTA, into shared memory
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < kRedRep; ++r) tot += __ldcg(red + (size_t)(1 + r) * 2 * C + i);
    s_acc[i] = tot;
    if (blockIdx.x == 0) red[i] = tot;   // for the caller: dbeta = red[0][c], dgamma = red[1][c]
  }
  __syncthreads();
  float k0[8], k1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = t.g * 8 + j;
    const float r0 = s_acc[c] * inv_count, r1 = s_acc[C + c] * inv_count;
    k1[j] = -sc[j] * r1 * rstd[c];
    k0[j] = -sc[j] * r0 - k1[j] * mean[c];
  }
  {
    BwdBatch<DY2, U> cur, nxt;
    bwd_load(cur, dy1, dy1_ld, dy2, dy2_ld, z, z_ld, last_base, npix, t);
    for (int64_t base = last_base; base >= 0; base -= step) {
      bwd_load(nxt, dy1, dy1_ld, dy2, dy2_ld, z, z_ld, base - step, npix, t);   // base - step < 0: nothing
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = base + u * t.py + t.ty;
        if (p < npix) {
          float d[8], zz[8], e[8], o[8];
          unpack8(cur.d[u], d);
          unpack8(cur.z[u], zz);
          if (DY2) {
            unpack8(cur.e[u], e);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] += e[j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gg = d[j] * act_grad(zz[j] * sc[j] + sh[j], act, slope);
            o[j] = sc[j] * gg + k1[j] * zz[j] + k0[j];
          }
          store8(dz + p * dz_ld + t.g * 8, o);
        }
      }
      cur = nxt;
    }
  }
}

// Biased conv + LeakyReLU layers (no BatchNorm): dz = (dy1+dy2) * leaky'(a) with a the stored
// activation (sign(a) == sign(pre-activation)); dbias[c] += sum dz.
__global__ void __launch_bounds__(256, 2)
act_bwd_bias_kernel(const __nv_bfloat16* __restrict__ dy1, int dy1_ld,
                    const __nv_bfloat16* __restrict__ dy2, int dy2_ld,
                    const __nv_bfloat16* __restrict__ a, int a_ld, __nv_bfloat16* __restrict__ dz,
                    int dz_ld, int C, int64_t npix, int act, float slope,
                    float* __restrict__ dbias) {
  extern __shared__ float s_acc[];  // [C]
  const Slot t = slot_of(C);
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float acc[8] = {0};
  const int64_t step = (int64_t)gridDim.x * t.py * UNR3;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int64_t base = (int64_t)blockIdx.x * t.py * UNR3; base < npix; base += step) {
    uint4 rd[UNR3], ra[UNR3], re[UNR3];
#pragma unroll
    for (int u = 0; u < UNR3; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      const bool ok = p < npix;
      rd[u] = ok ? ldraw(dy1 + p * dy1_ld + t.g * 8) : zero4;
      ra[u] = ok ? ldraw(a + p * a_ld + t.g * 8) : zero4;
      re[u] = (ok && dy2 != nullptr) ? ldraw(dy2 + p * dy2_ld + t.g * 8) : zero4;
    }
#pragma unroll
    for (int u = 0; u < UNR3; ++u) {
      const int64_t p = base + u * t.py + t.ty;
      if (p < npix) {
        float d[8], aa[8], e[8];
        unpack8(rd[u], d);
        unpack8(ra[u], aa);
        unpack8(re[u], e);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          d[j] = (d[j] + e[j]) * act_grad(aa[j], act, slope);
          acc[j] += d[j];
        }
        store8(dz + p * dz_ld + t.g * 8, d);
      }
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[t.g * 8 + j], acc[j]);
  }
  __syncthreads();
  if (dbias != nullptr)
    flush_add_v4(dbias, s_acc, C, threadIdx.x, blockDim.x);
}

// fp32 -> bf16 copy (gradient of the fp32 low-resolution logits entering the bf16 GEMMs)
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t count4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

static int reduce_grid(int64_t npix, int py, int cap = 148 * 8) {
  int64_t blocks = (npix + py * UNR3 - 1) / (py * UNR3);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int stream_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int check_c(int C, const char* what) {
  if (C % 8 || C < 8 || C > 2048) return set_error(B200_EINVAL, "%s: C=%d must be a multiple of 8 in [8,2048]", what, C);
  return B200_OK;
}
// threads per block so that blockDim is a multiple of the channel-group count
static int block_for(int C) {
  const int groups = C / 8;
  int t = (256 / groups) * groups;
  if (t == 0) t = groups;  // C = 2048 + : one row of groups
  return t;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_channel_stats(const void* x, int ld, int C, int64_t npix, float* stats, cudaStream_t stream) {
  int rc = check_c(C, "channel_stats");
  if (rc) return rc;
  const int threads = block_for(C);
  channel_stats_kernel<<<reduce_grid(npix, threads / (C / 8), 148 * 4), threads, 2 * C * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(x), ld, C, npix, stats);
  return check_launch("channel_stats");
}

int b200_bn_finalize(const float* stats, int C, float count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps,
                     int training, float* scale, float* shift, float* mean_out, float* rstd_out,
                     cudaStream_t stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, C, count, gamma, beta, running_mean,
                                                          running_var, momentum, eps, training,
                                                          scale, shift, mean_out, rstd_out);
  return check_launch("bn_finalize");
}

int b200_bn_act_apply(const void* x, int x_ld, void* y, int y_ld, int C, int64_t npix,
                      const float* scale, const float* shift, int act, float slope,
                      cudaStream_t stream) {
  int rc = check_c(C, "bn_act_apply");
  if (rc) return rc;
  const int threads = block_for(C);
  bn_act_apply_kernel<<<reduce_grid(npix, threads / (C / 8)), threads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<__nv_bfloat16*>(y), y_ld, C, npix,
      const_cast<float*>(scale), const_cast<float*>(shift), act, slope, nullptr, 1.f, nullptr, nullptr,
      nullptr, nullptr, 0.f, 0.f, 1, nullptr, nullptr, 0);
  return check_launch("bn_act_apply");
}

// BatchNorm finalisation fused into the normalise+activate pass (one launch per layer): see
// bn_act_apply_kernel.  scale/shift/mean_out/rstd_out receive the coefficients for the backward.
int b200_bn_norm_act(const void* x, int x_ld, void* y, int y_ld, int C, int64_t npix,
                     const float* stats, float count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps,
                     int training, int act, float slope, float* scale, float* shift,
                     float* mean_out, float* rstd_out, cudaStream_t stream) {
  int rc = check_c(C, "bn_norm_act");
  if (rc) return rc;
  if (training && stats == nullptr) return set_error(B200_EINVAL, "bn_norm_act: training needs stats");
  const int threads = block_for(C);
  bn_act_apply_kernel<<<reduce_grid(npix, threads / (C / 8)), threads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<__nv_bfloat16*>(y), y_ld, C, npix, scale,
      shift, act, slope, training ? stats : nullptr, count, gamma, beta, running_mean, running_var,
      momentum, eps, training, mean_out, rstd_out, 1);
  return check_launch("bn_norm_act");
}

int b200_bn_act_bwd_reduce(const void* dy1, int dy1_ld, const void* dy2, int dy2_ld, const void* z,
                           int z_ld, int C, int64_t npix, const float* scale, const float* shift,
                           const float* mean, const float* rstd, int act, float slope, float* red,
                           cudaStream_t stream) {
  int rc = check_c(C, "bn_act_bwd_reduce");
  if (rc) return rc;
  const int threads = block_for(C);
  bn_act_bwd_reduce_kernel<<<reduce_grid(npix, threads / (C / 8), 148 * 4), threads, 2 * C * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(dy1), dy1_ld, static_cast<const __nv_bfloat16*>(dy2), dy2_ld,
      static_cast<const __nv_bfloat16*>(z), z_ld, C, npix, scale, shift, mean, rstd, act, slope, red);
  return check_launch("bn_act_bwd_reduce");
}

int b200_bn_act_bwd_apply(const void* dy1, int dy1_ld, const void* dy2, int dy2_ld, const void* z,
                          int z_ld, void* dz, int dz_ld, int C, int64_t npix, const float* scale,
                          const float* shift, const float* mean, const float* rstd,
                          const float* red, float inv_count, int act, float slope,
                          cudaStream_t stream) {
  int rc = check_c(C, "bn_act_bwd_apply");
  if (rc) return rc;
  const int threads = block_for(C);
  bn_act_bwd_apply_kernel<<<reduce_grid(npix, threads / (C / 8)), threads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dy1), dy1_ld, static_cast<const __nv_bfloat16*>(dy2), dy2_ld,
      static_cast<const __nv_bfloat16*>(z), z_ld, static_cast<__nv_bfloat16*>(dz), dz_ld, C, npix,
      scale, shift, mean, rstd, red, inv_count, act, slope);
  return check_launch("bn_act_bwd_apply");
}

// b200_bn_act_bwd_reduce + b200_bn_act_bwd_apply in one cooperative launch; red must be zeroed.
int b200_bn_act_bwd_fused(const void* dy1, int dy1_ld, const void* dy2, int dy2_ld, const void* z,
                          int z_ld, void* dz, int dz_ld, int C, int64_t npix, const float* scale,
                          const float* shift, const float* mean, const float* rstd, float* red,
                          float inv_count, int act, float slope, cudaStream_t stream) {
  int rc = check_c(C, "bn_act_bwd_fused");
  if (rc) return rc;
  const int threads = block_for(C);
  const int py = threads / (C / 8);
  const size_t smem = 2 * C * sizeof(float);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  // two co-resident CTAs of <= 256 threads per SM (128 registers per thread): the whole grid must
  // be resident for the grid barrier
  // pixels per batch and thread: the deepest that still leaves every SM its two CTAs (a small layer would rather
  // have the CTAs: 4096 pixels x 128 channels are 64 CTAs at 4 pixels per batch, 256 at one)
  const int cap0 = n_sm * (threads <= 256 ? 2 : 1);
  auto blocks_at = [&](int u) { return (npix + (int64_t)py * u - 1) / ((int64_t)py * u); };
  const int unr = (dy2 == nullptr && blocks_at(UNRF1) >= cap0) ? UNRF1 : (blocks_at(UNRF) >= cap0 ? UNRF : 1);
  const int64_t blocks = blocks_at(unr);
  const int cap = n_sm * (threads <= 256 ? 2 : 1);
  const int grid = (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
  const __nv_bfloat16* a0 = static_cast<const __nv_bfloat16*>(dy1);
  const __nv_bfloat16* a1 = static_cast<const __nv_bfloat16*>(dy2);
  const __nv_bfloat16* a2 = static_cast<const __nv_bfloat16*>(z);
  __nv_bfloat16* a3 = static_cast<__nv_bfloat16*>(dz);
  void* args[] = {&a0, &dy1_ld, &a1, &dy2_ld, &a2, &z_ld, &a3, &dz_ld, &C, &npix, &scale, &shift,
                  &mean, &rstd, &red, &inv_count, &act, &slope};
  const void* fn = dy2 != nullptr ? (unr == 1 ? (const void*)bn_act_bwd_fused_kernel<true, 1> : (const void*)bn_act_bwd_fused_kernel<true, UNRF>)
                   : unr == UNRF1 ? (const void*)bn_act_bwd_fused_kernel<false, UNRF1>
                   : unr == UNRF  ? (const void*)bn_act_bwd_fused_kernel<false, UNRF>
                                  : (const void*)bn_act_bwd_fused_kernel<false, 1>;
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, stream);
  if (e != cudaSuccess) return set_error(B200_ECUDA, "bn_act_bwd_fused: %s", cudaGetErrorString(e));
  return check_launch("bn_act_bwd_fused");
}

int b200_cast_f32_bf16(const float* x, void* y, int64_t count, cudaStream_t stream) {
  if (count % 4) return set_error(B200_EINVAL, "cast_f32_bf16: count must be a multiple of 4");
  cast_f32_bf16_kernel<<<stream_grid(count / 4), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y), count / 4);
  return check_launch("cast_f32_bf16");
}

int b200_act_bwd_bias(const void* dy1, int dy1_ld, const void* dy2, int dy2_ld, const void* a,
                      int a_ld, void* dz, int dz_ld, int C, int64_t npix, int act, float slope,
                      float* dbias, cudaStream_t stream) {
  int rc = check_c(C, "act_bwd_bias");
  if (rc) return rc;
  const int threads = block_for(C);
  act_bwd_bias_kernel<<<reduce_grid(npix, threads / (C / 8), 148 * 4), threads, C * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(dy1), dy1_ld, static_cast<const __nv_bfloat16*>(dy2), dy2_ld,
      static_cast<const __nv_bfloat16*>(a), a_ld, static_cast<__nv_bfloat16*>(dz), dz_ld, C, npix,
      act, slope, dbias);
  return check_launch("act_bwd_bias");
}

}  // extern "C"
