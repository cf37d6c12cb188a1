"""Glue between ``nn.Module`` and the hand-written forward/backward of ops.py.

A module implements ``_fwd(*nhwc_inputs) -> (outputs, ctx)`` and ``_bwd(ctx, *douts) -> (dinputs,
{param: grad})``; :func:`run_module` exposes that pair to autograd as ONE node, so the module is a
drop-in for the reference class of the same name (logical NCHW tensors in and out, fp32 parameters
receiving ``.grad``).  Tensors that cross a module boundary are bf16 channels-last, i.e. NHWC in
memory; anything else is converted once at the boundary.
"""
import torch

from .. import _lib
from .. import kernels as K
from .. import ops

BF16 = torch.bfloat16
_pending_nbt = []


def bn_tuple(bn, training):
    """(gamma, beta, running_mean, running_var) of an nn.BatchNorm2d; queues the
    num_batches_tracked increment that nn.BatchNorm2d performs in training mode."""
    if training and bn.num_batches_tracked is not None:
        _pending_nbt.append(bn.num_batches_tracked)
    return (bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var)


def flush_nbt():
    if _pending_nbt:
        torch._foreach_add_(_pending_nbt, 1)
        del _pending_nbt[:]


def to_nhwc(x):
    """Logical NCHW tensor -> bf16 NHWC view/copy with unit channel stride and 16-byte pixels."""
    v = x.permute(0, 2, 3, 1)
    if (x.dtype == BF16 and v.stride(3) == 1 and v.stride(2) % 8 == 0
            and v.stride(1) == v.shape[2] * v.stride(2) and v.stride(0) == v.shape[1] * v.stride(1)
            and v.data_ptr() % 16 == 0):
        return v
    n, c, h, w = x.shape
    cpad = K.round_up(c, 8)
    buf = torch.empty((n, h, w, cpad), dtype=BF16, device=x.device)
    if cpad != c:
        buf[..., c:].zero_()
    buf[..., :c].copy_(v)
    return buf[..., :c]


def to_nchw(a):
    return a.permute(0, 3, 1, 2)


class ZeroBordered(torch.Tensor):
    """Marker type of ``losses.upsample_softmax(..., zero_border=True)``: a logical ``[N, C, H, W]`` bf16
    tensor that is the interior of a ``[N, H+2, W+2, 32]`` NHWC buffer whose one-pixel border and padding
    channels are ZERO (Conv2d's padding=1 materialised).  Views (``detach()``) keep the type and the strides;
    anything that allocates new memory loses the stride pattern, so type + strides together identify it."""


def padded_strides(x):
    """True when the logical NCHW tensor x has the strides of the interior of a [N, H+2, W+2, 32] buffer."""
    if x.dim() != 4 or x.dtype != BF16:
        return False
    n, c, h, w = x.shape
    row = (w + 2) * 32
    return (c <= 32 and h % 2 == 0 and w % 2 == 0 and tuple(x.stride()) == ((h + 2) * row, 1, row, 32)
            and x.storage_offset() >= row + 32 and x.data_ptr() % 64 == 0)


def padded_base(x):
    """The whole [N, H+2, W+2, 32] buffer behind a tensor with padded_strides."""
    n, c, h, w = x.shape
    row = (w + 2) * 32
    return torch.Tensor.as_strided(x.detach().as_subclass(torch.Tensor), (n, h + 2, w + 2, 32), ((h + 2) * row, row, 32, 1),
                                   x.storage_offset() - row - 32)


def zero_bordered_base(x):
    """-> the zero-bordered [N, H+2, W+2, 32] buffer when x is one (see ZeroBordered), else None."""
    if isinstance(x, ZeroBordered) and padded_strides(x):
        return padded_base(x)
    return None


def module_params(module):
    ps = getattr(module, "_b200_params", None)
    if ps is None:
        ps = list(module.parameters())
        object.__setattr__(module, "_b200_params", ps)
    return ps


class B200Module(torch.nn.Module):
    """Base of the drop-in modules: default NCHW<->NHWC boundary handling around _fwd / _bwd."""

    def _fwd_api(self, *inputs):
        outs, ctx = self._fwd(*[to_nhwc(x) for x in inputs])
        if isinstance(outs, tuple):
            return tuple(to_nchw(o) for o in outs), ctx
        return to_nchw(outs), ctx

    def _bwd_api(self, ctx, *douts, need_dx=True, need_dw=True):
        ds = [None if d is None else to_nhwc(d) for d in douts]
        dins, grads = self._bwd(ctx, *ds, need_dx=need_dx)
        if isinstance(dins, (tuple, list)):
            return tuple(None if d is None else to_nchw(d) for d in dins), grads
        return (None if dins is None else to_nchw(dins)), grads

    def forward(self, *inputs):
        return run_module(self, *inputs)


class _ModuleFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, n_inputs, *args):
        inputs = args[:n_inputs]
        ops.ARENA.begin((id(module), "f"), inputs[0].device)
        try:
            outs, saved = module._fwd_api(*inputs)
        finally:
            ops.ARENA.end()
        ctx.module, ctx.saved, ctx.n_inputs, ctx.params = module, saved, n_inputs, args[n_inputs:]
        ctx.single = not isinstance(outs, tuple)
        return outs

    @staticmethod
    def backward(ctx, *douts):
        needs_in = ctx.needs_input_grad[2:2 + ctx.n_inputs]
        needs_p = ctx.needs_input_grad[2 + ctx.n_inputs:]
        object.__setattr__(ctx.module, "_b200_dirty", True)  # an optimizer step is about to follow
        dev = next(d.device for d in douts if d is not None)
        ops.ARENA.begin((id(ctx.module), "b", tuple(d is not None for d in douts), any(needs_p)), dev)
        ops.WSCRATCH.begin((id(ctx.module), "b", tuple(d is not None for d in douts), any(needs_p)), dev)
        ops.REDUCER.begin(any(needs_p), dev, ctx.params)
        try:
            dins, grads = ctx.module._bwd_api(ctx.saved, *douts, need_dx=any(needs_in), need_dw=any(needs_p))
            ops.WSCRATCH.end()  # one launch: tap-major weight-gradient scratch -> PyTorch-layout gradients
        except BaseException:
            ops.WSCRATCH.end(ok=False)
            raise
        finally:
            ops.REDUCER.end()   # exchanges what is left of this backward's gradient span (world > 1)
            ops.ARENA.end()
        ctx.saved = None
        if not isinstance(dins, (tuple, list)):
            dins = (dins,)
        dins = tuple(d if need else None for d, need in zip(dins, needs_in))
        pg = tuple((grads.get(p) if need else None) for p, need in zip(ctx.params, needs_p))
        return (None, None) + dins + pg


def run_module(module, *inputs):
    dev = inputs[0].device
    if dev.type != "cuda":
        raise _lib.B200Error("%s only runs on a B200 (got a %s tensor); there is no CPU fallback"
                             % (type(module).__name__, dev.type))
    _lib.ensure_device(dev.index or 0)
    _lib.set_device_index(dev.index or 0)
    ops.refresh_packs(module, force=module.training or getattr(module, "_b200_dirty", False))
    object.__setattr__(module, "_b200_dirty", False)
    out = _ModuleFunction.apply(module, len(inputs), *inputs, *module_params(module))
    flush_nbt()
    return out
