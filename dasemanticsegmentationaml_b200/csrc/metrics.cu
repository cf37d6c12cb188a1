// Evaluation metrics on the device: the 19x19 confusion matrix ("fast_hist") and argmax.
//   reference: utils.py:161-167 (fast_hist), utils.py:98-122 (reverse_one_hot), utils.py:151-159
#include <stdint.h>

#include "ptx.cuh"
#include "status.h"
#include "b200seg.h"

namespace b200 {

constexpr int kHistThreads = 256;
constexpr int kMaxBins = 1024;  // n <= 32 classes

// hist[n*a + b] += 1 for every pixel with 0 <= a < n (a = label, b = prediction).
// Each warp owns a private shared-memory copy of the n*n bins (32-bit counts are safe: a CTA
// sees at most 2^31 pixels by construction of the grid-stride loop), merged into the int64
// global matrix with one atomic per non-empty bin per CTA.
// A prediction outside [0, n) lands in bin n*a+b exactly as np.bincount would place it as long as
// the index stays below n*n; indices >= n*n (numpy would grow the array and the reshape in
// utils.py:167 would then raise) are reported through *bad.
template <typename LabelT, typename PredT>
__global__ void __launch_bounds__(kHistThreads)
fast_hist_kernel(const LabelT* __restrict__ a, const PredT* __restrict__ b, int64_t count, int n,
                 unsigned long long* __restrict__ hist, int* __restrict__ bad) {
  extern __shared__ uint32_t s_bins[];  // [warps][n*n]
  const int nn = n * n;
  const int warps = kHistThreads / 32;
  for (int i = threadIdx.x; i < warps * nn; i += kHistThreads) s_bins[i] = 0;
  __syncthreads();
  uint32_t* mine = s_bins + (threadIdx.x >> 5) * nn;
  const int64_t stride = (int64_t)gridDim.x * kHistThreads;
  for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < count; i += stride) {
    const long long la = (long long)a[i];
    if (la >= 0 && la < n) {
      const long long idx = la * n + (long long)b[i];
      if (idx >= 0 && idx < nn)
        atomicAdd(&mine[idx], 1u);
      else
        atomicExch(bad, 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nn; i += kHistThreads) {
    unsigned long long tot = 0;
    for (int w = 0; w < warps; ++w) tot += s_bins[w * nn + i];
    if (tot) atomicAdd(&hist[i], tot);
  }
}

// Number of positions where pred == label (compute_global_accuracy's numerator, utils.py:151-159).
template <typename LabelT, typename PredT>
__global__ void __launch_bounds__(256)
count_equal_kernel(const LabelT* __restrict__ a, const PredT* __restrict__ b, int64_t count,
                   unsigned long long* __restrict__ out) {
  unsigned long long local = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    local += ((long long)a[i] == (long long)b[i]) ? 1ull : 0ull;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// Same per image: out[blockIdx.y] += #(pred == label) over that image's `per_image` positions (the
// evaluation loop keeps one counter per image so that the host reproduces the reference's
// mean of per-image ratios, train.py:50,56, with ONE device->host copy per evaluation).
template <typename LabelT, typename PredT>
__global__ void __launch_bounds__(256)
count_equal_batched_kernel(const LabelT* __restrict__ a, const PredT* __restrict__ b, int64_t per_image,
                           unsigned long long* __restrict__ out) {
  const LabelT* ai = a + (int64_t)blockIdx.y * per_image;
  const PredT* bi = b + (int64_t)blockIdx.y * per_image;
  unsigned long long local = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += stride)
    local += ((long long)ai[i] == (long long)bi[i]) ? 1ull : 0ull;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out + blockIdx.y, local);
}

static int hist_grid(int64_t count) {
  int64_t blocks = (count + kHistThreads * 16 - 1) / (kHistThreads * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace b200

using namespace b200;

extern "C" {

// label/pred: device arrays of `count` elements; elem_bytes 8 (int64, the reference's dtype),
// 4 (int32) or 1 (uint8, pred only as produced by the fused argmax). hist: int64[n*n], accumulated.
// bad: device int set to 1 when an index falls outside the n*n bins.
int b200_fast_hist(const void* label, int label_bytes, const void* pred, int pred_bytes,
                   int64_t count, int n, int64_t* hist, int* bad, cudaStream_t stream) {
  if (n < 1 || n * n > kMaxBins) return set_error(B200_EINVAL, "fast_hist: n=%d unsupported (n*n <= %d)", n, kMaxBins);
  if (count < 0) return set_error(B200_EINVAL, "fast_hist: negative count");
  if (count == 0) return B200_OK;
  const size_t smem = (size_t)(kHistThreads / 32) * n * n * sizeof(uint32_t);
  const int grid = hist_grid(count);
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
#define LAUNCH(LT, PT)                                                                         \
  fast_hist_kernel<LT, PT><<<grid, kHistThreads, smem, stream>>>(                              \
      static_cast<const LT*>(label), static_cast<const PT*>(pred), count, n, h, bad)
  if (label_bytes == 8 && pred_bytes == 8) LAUNCH(int64_t, int64_t);
  else if (label_bytes == 8 && pred_bytes == 1) LAUNCH(int64_t, uint8_t);
  else if (label_bytes == 8 && pred_bytes == 4) LAUNCH(int64_t, int32_t);
  else if (label_bytes == 4 && pred_bytes == 4) LAUNCH(int32_t, int32_t);
  else if (label_bytes == 1 && pred_bytes == 1) LAUNCH(uint8_t, uint8_t);
  else return set_error(B200_EINVAL, "fast_hist: unsupported element sizes %d/%d", label_bytes, pred_bytes);
#undef LAUNCH
  return check_launch("fast_hist");
}

int b200_count_equal(const void* label, int label_bytes, const void* pred, int pred_bytes,
                     int64_t count, int64_t* out, cudaStream_t stream) {
  if (count <= 0) return B200_OK;
  const int grid = hist_grid(count);
  unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
  if (label_bytes == 8 && pred_bytes == 8)
    count_equal_kernel<int64_t, int64_t><<<grid, 256, 0, stream>>>(static_cast<const int64_t*>(label), static_cast<const int64_t*>(pred), count, o);
  else if (label_bytes == 8 && pred_bytes == 1)
    count_equal_kernel<int64_t, uint8_t><<<grid, 256, 0, stream>>>(static_cast<const int64_t*>(label), static_cast<const uint8_t*>(pred), count, o);
  else
    return set_error(B200_EINVAL, "count_equal: unsupported element sizes %d/%d", label_bytes, pred_bytes);
  return check_launch("count_equal");
}

int b200_count_equal_batched(const void* label, int label_bytes, const void* pred, int pred_bytes,
                             int n_images, int64_t per_image, int64_t* out, cudaStream_t stream) {
  if (n_images <= 0 || per_image <= 0) return B200_OK;
  if (n_images > 65535) return set_error(B200_EINVAL, "count_equal_batched: at most 65535 images per call");
  int gx = hist_grid(per_image);
  if (gx > 148 * 2) gx = 148 * 2;
  const dim3 grid(gx, n_images);
  unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
  if (label_bytes == 8 && pred_bytes == 8)
    count_equal_batched_kernel<int64_t, int64_t><<<grid, 256, 0, stream>>>(static_cast<const int64_t*>(label), static_cast<const int64_t*>(pred), per_image, o);
  else if (label_bytes == 8 && pred_bytes == 1)
    count_equal_batched_kernel<int64_t, uint8_t><<<grid, 256, 0, stream>>>(static_cast<const int64_t*>(label), static_cast<const uint8_t*>(pred), per_image, o);
  else if (label_bytes == 1 && pred_bytes == 1)
    count_equal_batched_kernel<uint8_t, uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(label), static_cast<const uint8_t*>(pred), per_image, o);
  else
    return set_error(B200_EINVAL, "count_equal_batched: unsupported element sizes %d/%d", label_bytes, pred_bytes);
  return check_launch("count_equal_batched");
}

}  // extern "C"
