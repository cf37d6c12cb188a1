// Parameter blocks shared by the tcgen05 implicit-GEMM convolution kernels and their host
// launchers (igemm.cu).  See DESIGN.md "Dense convolutions" for the data layout.
#pragma once
#include <stdint.h>

namespace b200 {

constexpr int kMaxTaps = 16;    // 4x4 discriminator filters
constexpr int kMaxClasses = 4;  // output-parity classes of a stride-2 dgrad

// One launch = one implicit GEMM  out[pixel, co] = sum_{tap, ci} in[pixel*stride + tap, ci] * w[co, tap, ci].
// A "class" (blockIdx.z) is a sub-grid of output pixels (used by strided dgrad, where the four
// output parities see different filter taps); a forward conv has exactly one class.
struct ConvParams {
  // output sub-grid iterated by patch tiles, per class
  int Ho[kMaxClasses], Wo[kMaxClasses];
  int tiles_h[kMaxClasses], tiles_w[kMaxClasses];
  int n_taps[kMaxClasses];
  int8_t tap_dh[kMaxClasses][kMaxTaps];
  int16_t tap_dw[kMaxClasses][kMaxTaps];  // (16 bits: the zero-bordered pair view of the probabilities puts its second row a map width away)
  int8_t tap_k[kMaxClasses][kMaxTaps];  // index of the tap's K-slab in the packed filter
  int oa[kMaxClasses], ob[kMaxClasses];  // output pixel offset of the class
  int n_img;
  int th, tw;        // sub-tile patch: th*tw == 128 output pixels (one M=128 accumulator)
  int mt;            // sub-tiles per CTA, stacked along h: the CTA covers (th*mt) x tw pixels and
                     // reuses each filter tile for mt accumulators (less L2 traffic per FLOP)
  int in_stride;     // traversal stride over the input
  int cin_pad;       // K elements per tap in the packed filter (multiple of KC)
  int KC;            // channels per pipeline stage: 64 (128B swizzle) or 32 (64B swizzle)
  int BN;            // output channels per CTA (16..256, multiple of 16)
  int stages;
  // epilogue
  void* out;
  int out_ld;        // channel stride of an output pixel (elements)
  int out_coff;      // first channel written
  int Hout, Wout;    // full output map
  int os;            // output pixel stride (2 for strided dgrad, else 1)
  const float* bias; // nullptr or [>= grid.y*BN]
  int act;           // 0 none, 1 relu, 2 leaky relu
  float slope;
  int out_f32;       // 1: fp32 output, 0: bf16
  float* stats;      // nullptr, or [2][C] per-channel sum / sum of squares of the fp32 results
  int stats_ld;      // C
  // Backward of a biased conv + LeakyReLU layer fused into the data-gradient launch that produces
  // its output gradient: out = acc * (mask[pixel, c] > 0 ? 1 : mask_slope), where `mask` is the
  // layer's stored activation (sign(a) == sign(pre-activation)); with stats_sum_only the statistics
  // path then yields stats[0][c] = sum of the masked gradient = the bias gradient.
  const void* mask;  // nullptr, or bf16 NHWC view shaped like the output
  int mask_ld;
  float mask_slope;
  int stats_sum_only;
  int wide;          // 1: output (and mask) rows are 32-byte aligned -> 256-bit stores / loads
  int kg;            // K-blocks (of KC channels) per pipeline stage (CTA-pair kernel)
  // Input-halo reuse (CTA-pair kernel): per K-block the activations are fetched ONCE as `np` boxes of
  // box_h x box_w pixels (one box for a stride-1 conv; four input-parity planes, fetched with TMA element
  // stride 2, for a stride-2 conv) and every tap is a shifted UMMA descriptor window into one of them.
  int halo;          // 0 off, 1 on
  int np;            // boxes (planes) per K-block: 1 or 4
  int box_h, box_w;  // pixels per box (th + extent, tw + extent)
  int pl_dh[4], pl_dw[4];                 // input pixel of a box's first element, relative to (h0, w0) * in_stride
  int8_t tap_pl[kMaxClasses][kMaxTaps];   // plane of a tap
  int8_t tap_off[kMaxClasses][kMaxTaps];  // first row of the tap's window inside its box (oh * box_w + ow)
  int tb, sb;        // taps per filter-ring slot, filter-ring slots
  int bres;          // 1: the pair's whole filter tile (all slabs, all K-blocks) stays resident in shared memory
  int tma_out;       // 1: bf16 output tiles leave through TMA stores (CTA-pair kernel; one tensor map per class)
  // diagnostic (b200_debug_timeline): CTA pair 0 records %globaltimer at its phase boundaries here
  //   [0] kernel entry  [1] set-up done  [2] first operands landed  [3] exit
  //   [8 + 4*lt + {0: MMAs of tile lt start, 1: committed, 2: epilogue starts (accumulator full), 3: epilogue done}]
  long long* dbg;
  int dbg_knob;      // diagnostic (b200_debug_knob): epilogue parts switched off for timing experiments (WRONG results)
};

// dW[co, ci, tap] += sum_pixels dz[pixel, co] * x[pixel*stride + tap, ci]
struct WgradParams {
  int Ho, Wo, tiles_h, tiles_w, n_img;
  int th, tw;       // th*tw == KPIX pixels (the GEMM K extent) per stage
  int in_stride;
  int n_taps;
  int8_t tap_dh[kMaxTaps];
  int16_t tap_dw[kMaxTaps];
  int RS;           // taps in the PyTorch filter (dW is [Cout][Cin][RS])
  int8_t tap_rs[kMaxTaps];
  int Cout, Cin;
  int co_tiles, ci_tiles;
  int BNW;          // ci per CTA: 64 / 128 / 256
  int stages;
  int splits;       // pixel-range splits (gridDim.x)
  float* dw;        // fp32, PyTorch layout, accumulated atomically
  // Optional tap-major scratch accumulator [RS][Cout][ci_pad] (fp32, zeroed): a thread's 16
  // consecutive input channels are contiguous there, so the split-K partials are added with four
  // 16-byte vector reductions instead of 16 scalar atomics 4*RS bytes apart (the scalar atomics of the
  // 3x3 / 4x4 layers cost 0.8 ms of a 15.3 ms step); b200_wgrad_unscratch permutes it into dw.
  float* scratch;
  int ci_pad;
  // all-taps kernel with input-halo reuse (see ConvParams::halo): np > 0 -> the input is fetched as np
  // boxes of box_h x box_w pixels per pixel tile and tap t is a shifted window into box tap_pl[t]
  int np, box_h, box_w;
  int pl_dh[4], pl_dw[4];
  int8_t tap_pl[kMaxTaps];
  int8_t tap_off[kMaxTaps];
};

// Decomposition of a tap list into input-parity planes and per-tap window offsets (igemm.cu).
struct TapPlanes {
  int np, eh, ew;
  int pl_dh[4], pl_dw[4];
  int8_t tap_pl[kMaxTaps];
  int8_t tap_qh[kMaxTaps], tap_qw[kMaxTaps];   // window offset inside the plane, in pixels
};

}  // namespace b200
