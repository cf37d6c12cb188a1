"""B200-native (sm_100a) implementation of the STDC-BiSeNet + adversarial-discriminator train /
eval hot path of TiloccaS/DASemanticSegmentationAML, behind the reference's own module API.

The compute lives in libb200seg.so (hand-written CUDA: tcgen05/TMEM/TMA implicit-GEMM convolutions
and fused HBM-bound kernels, C ABI in include/b200seg.h); this package is the thin PyTorch host side:
drop-in ``nn.Module`` classes with the reference's names, constructor signatures and state_dict
layout, the loss composition of ``train.py`` and the mIoU metric.  There is no CPU fallback.
"""
__version__ = "0.1.0"
