// Multi-tensor optimizer updates: the four optimizer steps of a DA iteration
// (train.py:219-221, 235-237, 252-254, 260-262 -- SGD(momentum 0.9, wd 5e-4) on the segmentation net,
// Adam(betas 0.9 / 0.99) on the discriminator) each become ONE launch over all parameter tensors.
// Arithmetic follows torch.optim.SGD / torch.optim.Adam (L2 weight decay folded into the gradient,
// bias-corrected Adam, no amsgrad, no maximize).  HBM-bound: SGD moves 20 B per parameter (p, g, buf
// read; p, buf written), Adam 28 B.
//
// table: int64 [n_tensors][8] = {param, grad, state1 (momentum buffer / exp_avg), state2 (exp_avg_sq),
//        numel, first chunk index, hyper-parameter group, flags (bit 0: state1 is initialised)}
// hyper: float [n_groups][8] = {lr, momentum, dampening, weight_decay, nesterov, beta1, beta2, eps}
// The table and the hyper-parameters live in device memory so that a captured CUDA graph picks up a
// new learning rate (poly_lr_scheduler) or refreshed pointers without re-capture.
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

constexpr int kOptChunk = 4096;   // elements per CTA
constexpr int kOptThreads = 256;

struct OptSlot {
  float* p;
  const float* g;
  float* s1;
  float* s2;
  int64_t numel, first_chunk;
  int group, flags;
};

__device__ __forceinline__ OptSlot find_slot(const long long* __restrict__ table, int n, int chunk, int* local_chunk) {
  int lo = 0, hi = n - 1;   // last tensor whose first chunk <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[(size_t)mid * 8 + 5] <= chunk) lo = mid;
    else hi = mid - 1;
  }
  const long long* t = table + (size_t)lo * 8;
  OptSlot s;
  s.p = reinterpret_cast<float*>(t[0]);
  s.g = reinterpret_cast<const float*>(t[1]);
  s.s1 = reinterpret_cast<float*>(t[2]);
  s.s2 = reinterpret_cast<float*>(t[3]);
  s.numel = t[4];
  s.first_chunk = t[5];
  s.group = (int)t[6];
  s.flags = (int)t[7];
  *local_chunk = chunk - (int)s.first_chunk;
  return s;
}

// torch.optim.SGD: g += wd * p; buf = first ? g : momentum * buf + (1 - dampening) * g;
//                  g = nesterov ? g + momentum * buf : buf; p -= lr * g
__global__ void __launch_bounds__(kOptThreads)
sgd_multi_kernel(const long long* __restrict__ table, int n, const float* __restrict__ hyper) {
  int lc;
  const OptSlot s = find_slot(table, n, blockIdx.x, &lc);
  const float* h = hyper + s.group * 8;
  const float lr = h[0], mom = h[1], damp = h[2], wd = h[3];
  const bool nesterov = h[4] != 0.f, has_buf = s.s1 != nullptr, init = (s.flags & 1) != 0;
  const int64_t begin = (int64_t)lc * kOptChunk;
  const int64_t end = begin + kOptChunk < s.numel ? begin + kOptChunk : s.numel;
  const bool vec = ((reinterpret_cast<uintptr_t>(s.p) | reinterpret_cast<uintptr_t>(s.g) |
                     reinterpret_cast<uintptr_t>(s.s1)) & 15) == 0 && (end - begin) % 4 == 0;
  if (vec) {
    for (int64_t i = begin + threadIdx.x * 4; i < end; i += kOptThreads * 4) {
      float4 p = *reinterpret_cast<const float4*>(s.p + i);
      const float4 g4 = *reinterpret_cast<const float4*>(s.g + i);
      float4 b = has_buf && init ? *reinterpret_cast<const float4*>(s.s1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float pv[4] = {p.x, p.y, p.z, p.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float g = gv[k] + wd * pv[k];
        if (has_buf) {
          bv[k] = init ? mom * bv[k] + (1.f - damp) * g : g;
          g = nesterov ? g + mom * bv[k] : bv[k];
        }
        pv[k] -= lr * g;
      }
      *reinterpret_cast<float4*>(s.p + i) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      if (has_buf) *reinterpret_cast<float4*>(s.s1 + i) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    }
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += kOptThreads) {
      const float p = s.p[i];
      float g = s.g[i] + wd * p;
      if (has_buf) {
        const float b = init ? mom * s.s1[i] + (1.f - damp) * g : g;
        s.s1[i] = b;
        g = nesterov ? g + mom * b : b;
      }
      s.p[i] = p - lr * g;
    }
  }
}

// torch.optim.Adam (capturable form): t = step + 1; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
// counters: int64[2] = {step, finished CTAs}: every CTA reads `step` on entry; the last CTA to finish
// publishes step + 1 and clears the ticket, so one launch is one optimizer step under graph replay.
__global__ void __launch_bounds__(kOptThreads)
adam_multi_kernel(const long long* __restrict__ table, int n, const float* __restrict__ hyper,
                  long long* __restrict__ counters) {
  __shared__ float s_bc[2];
  int lc;
  const OptSlot s = find_slot(table, n, blockIdx.x, &lc);
  const float* h = hyper + s.group * 8;
  const long long t = counters[0] + 1;
  if (threadIdx.x == 0) {
    s_bc[0] = (float)(1.0 - pow((double)h[5], (double)t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)h[6], (double)t));
  }
  __syncthreads();
  const float lr = h[0], wd = h[3], b1 = h[5], b2 = h[6], eps = h[7];
  const float step_size = lr / s_bc[0], bc2_sqrt = s_bc[1];
  const int64_t begin = (int64_t)lc * kOptChunk;
  const int64_t end = begin + kOptChunk < s.numel ? begin + kOptChunk : s.numel;
  for (int64_t i = begin + threadIdx.x; i < end; i += kOptThreads) {
    const float p = s.p[i];
    const float g = s.g[i] + wd * p;
    const float m = b1 * s.s1[i] + (1.f - b1) * g;
    const float v = b2 * s.s2[i] + (1.f - b2) * g * g;
    s.s1[i] = m;
    s.s2[i] = v;
    s.p[i] = p - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(counters + 1), 1ull);
    if (ticket == (unsigned long long)gridDim.x - 1) {
      counters[0] = t;
      counters[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_sgd_step(const int64_t* table_dev, int n_tensors, int total_chunks, const float* hyper_dev,
                  cudaStream_t stream) {
  if (n_tensors <= 0 || total_chunks <= 0) return B200_OK;
  sgd_multi_kernel<<<total_chunks, kOptThreads, 0, stream>>>(reinterpret_cast<const long long*>(table_dev), n_tensors, hyper_dev);
  return check_launch("sgd_step");
}

int b200_adam_step(const int64_t* table_dev, int n_tensors, int total_chunks, const float* hyper_dev,
                   int64_t* counters_dev, cudaStream_t stream) {
  if (n_tensors <= 0 || total_chunks <= 0) return B200_OK;
  adam_multi_kernel<<<total_chunks, kOptThreads, 0, stream>>>(reinterpret_cast<const long long*>(table_dev), n_tensors, hyper_dev,
                                                             reinterpret_cast<long long*>(counters_dev));
  return check_launch("adam_step");
}

int b200_optim_chunk(void) { return kOptChunk; }

}  // extern "C"
