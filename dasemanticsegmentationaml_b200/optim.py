"""Multi-tensor SGD / Adam: one kernel launch per ``optimizer.step()``.

Drop-ins for the optimizers the reference builds at train.py:170-172
(``torch.optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=5e-4)`` and
``torch.optim.Adam(model_D1.parameters(), lr, betas=(0.9, 0.99))``): same constructor arguments, same
arithmetic, the same ``state`` entries (``momentum_buffer``; ``step`` / ``exp_avg`` / ``exp_avg_sq``),
parameters stay fp32 ``nn.Parameter``s.  The pointer table and the hyper-parameters live in device
memory (the table is refreshed through a ring of pinned staging buffers), so ``step()`` is
safe inside CUDA-graph capture (tables created during a capture are uploaded by ``flush_pending()``
right after it -- ``train.GraphedStep`` does that); after changing ``param_groups[i]['lr']`` (``poly_lr_scheduler``)
call ``refresh_hyperparameters()`` before the next replay (``step()`` does it itself when it runs).
"""
import torch

from . import _lib
from . import kernels as K

_ROW = 8
_PENDING = []   # (host table, device table) pairs created during a graph capture, not yet uploaded


def flush_pending():
    """Upload the pointer tables that were created while a CUDA graph was being captured.  Must run
    after the capture has ended and before the first replay (train.GraphedStep does)."""
    while _PENDING:
        host, dev = _PENDING.pop()
        dev.copy_(host)


class _MultiTensorOptimizer(torch.optim.Optimizer):
    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self._key = None
        self._n = self._chunks = 0
        self._tab_dev = self._table = None
        self._staging, self._staging_next = [], 0
        self._captured, self._capture_pool = [], []
        self._hyp_dev = None
        self._hyp_values = None
        self._chunk = None
        self._n_elems = 0

    # -- hyper-parameters ---------------------------------------------------------------------
    def _hyper_row(self, group):
        raise NotImplementedError

    def refresh_hyperparameters(self, device=None):
        rows = [self._hyper_row(g) for g in self.param_groups]
        if self._hyp_dev is None:
            if device is None:
                return
            self._hyp_dev = torch.zeros((len(rows), _ROW), dtype=torch.float32, device=device)
        if rows != self._hyp_values:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.B200Error("hyper-parameters changed inside a CUDA-graph capture; call "
                                     "refresh_hyperparameters() before capturing / replaying")
            self._hyp_dev.copy_(torch.tensor(rows, dtype=torch.float32))   # rare: a blocking copy is fine
            self._hyp_values = rows

    # -- pointer table ------------------------------------------------------------------------
    def _state_tensors(self, p, group):
        raise NotImplementedError

    def _prepare(self):
        if self._chunk is None:
            self._chunk = int(_lib.lib().b200_optim_chunk())
        entries = []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise _lib.B200Error("the fused optimizers update fp32 CUDA parameters (got %s on %s)" % (p.dtype, p.device))
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                    raise _lib.B200Error("the fused optimizers need dense, contiguous fp32 gradients")
                s1, s2, flags = self._state_tensors(p, group)
                entries.append((p, g, s1, s2, gi, flags))
        if not entries:
            return False
        device = entries[0][0].device
        key = tuple((p.data_ptr(), g.data_ptr(), 0 if s1 is None else s1.data_ptr(), fl) for p, g, s1, _, _, fl in entries)
        if key != self._key:
            rows, chunk0 = [], 0
            for p, g, s1, s2, gi, flags in entries:
                rows.append([p.data_ptr(), g.data_ptr(), 0 if s1 is None else s1.data_ptr(),
                             0 if s2 is None else s2.data_ptr(), p.numel(), chunk0, gi, flags])
                chunk0 += (p.numel() + self._chunk - 1) // self._chunk
            n = len(rows)
            if torch.cuda.is_current_stream_capturing():
                # A table that is baked into a graph gets a device buffer of its own (an eager step in
                # between, with other gradient addresses, must not leak into the graph).  The buffer
                # must come from OUTSIDE the graph's memory pool -- memory allocated mid-capture
                # aliases earlier temporaries of the same graph, which would overwrite it on every
                # replay -- so it is taken from a few tables set aside by the eager warm-up steps.
                # Its content is constant for the life of the graph and is written ONCE after the
                # capture (flush_pending(), called by train.GraphedStep), not by a copy node per replay.
                if not self._capture_pool or self._capture_pool[-1].shape[0] < n:
                    raise _lib.B200Error("optimizer.step() inside a CUDA-graph capture needs an eager warm-up "
                                         "step first (and at most 4 distinct gradient sets per graph)")
                dev = self._capture_pool.pop()[:n]
                _PENDING.append((torch.tensor(rows, dtype=torch.int64), dev))
                self._captured.append(dev)
            else:
                # Eager steps: the host runs ahead of the GPU, so a pinned staging buffer may only be
                # rewritten once the asynchronous copy that last read it has executed -> ring of
                # staging buffers guarded by events (the device table itself is stream-ordered).
                cap = max(n, sum(len(g["params"]) for g in self.param_groups))
                if self._tab_dev is None or self._tab_dev.shape[0] < cap:
                    self._tab_dev = torch.zeros((cap, _ROW), dtype=torch.int64, device=device)
                    self._capture_pool = [torch.zeros((cap, _ROW), dtype=torch.int64, device=device) for _ in range(4)]
                    self._staging = [[torch.zeros((cap, _ROW), dtype=torch.int64).pin_memory(), None] for _ in range(4)]
                slot = self._staging[self._staging_next % len(self._staging)]
                self._staging_next += 1
                if slot[1] is not None:
                    slot[1].synchronize()
                slot[0][:n].copy_(torch.tensor(rows, dtype=torch.int64))
                dev = self._tab_dev[:n]
                dev.copy_(slot[0][:n], non_blocking=True)
                slot[1] = torch.cuda.Event()
                slot[1].record()
            self._table = dev
            self._key, self._n, self._chunks = key, n, chunk0
            self._n_elems = sum(e[0].numel() for e in entries)
        self.refresh_hyperparameters(device)
        _lib.set_device_index(device.index or 0)
        return True


class FusedSGD(_MultiTensorOptimizer):
    """``torch.optim.SGD`` (train.py:170) as one launch per step."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self._stepped = set()
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    def _hyper_row(self, g):
        return [float(g["lr"]), float(g["momentum"]), float(g["dampening"]), float(g["weight_decay"]),
                1.0 if g["nesterov"] else 0.0, 0.0, 0.0, 0.0]

    def _state_tensors(self, p, group):
        if group["momentum"] == 0:
            return None, None, 0
        st = self.state[p]
        if st.get("momentum_buffer") is None:
            st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
            self._stepped.discard(id(p))
        # torch: the first step copies the gradient into the buffer (flag 0), later steps blend
        return st["momentum_buffer"], None, 1 if id(p) in self._stepped else 0

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._stepped = {id(p) for p, st in self.state.items() if st.get("momentum_buffer") is not None}
        self._key = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._prepare():
            K.call("b200_sgd_step", K.ptr(self._table), K.c_int(self._n), K.c_int(self._chunks),
                   K.ptr(self._hyp_dev), K.stream(), nbytes=20.0 * self._n_elems, tag="%d tensors" % self._n)
            for group in self.param_groups:
                if group["momentum"] != 0:
                    self._stepped.update(id(p) for p in group["params"] if p.grad is not None)
        return loss


class FusedAdam(_MultiTensorOptimizer):
    """``torch.optim.Adam`` (train.py:172) as one launch per step; the step count lives on the device."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._counters = None

    def _hyper_row(self, g):
        return [float(g["lr"]), 0.0, 0.0, float(g["weight_decay"]), 0.0, float(g["betas"][0]), float(g["betas"][1]),
                float(g["eps"])]

    def _state_tensors(self, p, group):
        st = self.state[p]
        if "exp_avg" not in st:
            if self._counters is None:
                self._counters = torch.zeros(2, dtype=torch.int64, device=p.device)
            st["step"] = self._counters[0]          # one shared device counter, exposed per parameter
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st["exp_avg"], st["exp_avg_sq"], 1

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [st["step"] for st in self.state.values() if "step" in st]
        if steps:
            dev = next(st["exp_avg"].device for st in self.state.values() if "exp_avg" in st)
            self._counters = torch.zeros(2, dtype=torch.int64, device=dev)
            self._counters[0] = int(steps[0])
            for st in self.state.values():
                if "step" in st:
                    st["step"] = self._counters[0]
        self._key = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._prepare():
            K.call("b200_adam_step", K.ptr(self._table), K.c_int(self._n), K.c_int(self._chunks),
                   K.ptr(self._hyp_dev), K.ptr(self._counters), K.stream(), nbytes=28.0 * self._n_elems,
                   tag="%d tensors" % self._n)
        return loss
