"""Forward dispatch of ``models.MobileNetSkipAdd`` onto the C-ABI (one ``fd_forward`` per call).

Replaces the Python layer loop of reference models.py:706-732 and every PyTorch operator it
launches.  The engine is created lazily on the first forward so that
modules restored by ``torch.load`` without ``__init__`` (reference main.py:49-57) work.
"""
import torch

from . import plan as _plan

_SUPPORTED = (torch.float32, torch.float16, torch.bfloat16)


class SkipAddEngine:
    def __init__(self, module):
        self.module = module
        self.plans = {}          # (device index, n, h, w, dtype) -> Plan
        self.signature = None
        self.options = {}

    # -- weight freshness ----------------------------------------------------------------------
    def _weights_signature(self):
        """Per-call freshness check over EVERY parameter and buffer of the module: storage pointer, dtype and the
        autograd version counter (bumped by every in-place write through the tensor itself: ``load_state_dict``,
        ``p.copy_`` / ``p.mul_`` under ``no_grad``, optimizer steps).  ~250 tensors, a few tens of microseconds; catches
        .cuda()/.half()/.to() too.  Writes that bypass the counter (``p.data.mul_()``, a raw pointer handed to another
        library) still need an explicit ``refresh()``."""
        sig = []
        for t in self.module.parameters():
            sig.append((t.data_ptr(), t.dtype, t._version))
        for t in self.module.buffers():
            sig.append((t.data_ptr(), t.dtype, t._version))
        return tuple(sig)

    def refresh(self):
        """Drop packed weights (call after editing parameters in place)."""
        for p in self.plans.values():
            p.close()
        self.plans.clear()
        self.signature = None

    def set_option(self, name, value):
        """Forwarded to ``fd_plan_set_option`` on every current and future plan."""
        prev = self.options.get(name)
        self.options[name] = int(value)
        try:
            for p in self.plans.values():
                p.set_option(name, value)
        except Exception:
            if prev is None:
                self.options.pop(name, None)
            else:
                self.options[name] = prev
            raise

    # -- forward ---------------------------------------------------------------------------------
    def plan_for(self, x):
        m = self.module
        if m.training:
            raise RuntimeError("fastdepth_b200 is inference-only: call model.eval() first "
                               "(BatchNorm is folded from running statistics, reference main.py:65)")
        if not x.is_cuda:
            raise RuntimeError("fastdepth_b200: the accelerated forward needs a CUDA tensor; "
                               "there is no CPU fallback (use models.MobileNet for CPU plumbing)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError("expected input [N,3,H,W], got %s" % (tuple(x.shape),))
        wdtype = _plan._blocks_of(m)[0][0][0].weight.dtype
        if x.dtype != wdtype:
            raise RuntimeError("Input type (%s) and weight type (%s) should be the same" % (x.dtype, wdtype))
        if x.dtype not in _SUPPORTED:
            raise RuntimeError("unsupported dtype %s" % x.dtype)
        n, _, h, w = x.shape
        if h % 32 or w % 32:
            raise RuntimeError("The size of tensor a must match the size of tensor b: H and W must be multiples "
                               "of 32 for the skip connections to line up, got %dx%d" % (h, w))
        sig = self._weights_signature()
        if sig != self.signature:
            self.refresh()
            self.signature = sig
        key = (x.device.index, n, h, w, x.dtype)
        p = self.plans.get(key)
        if p is None:
            p = _plan.Plan.from_module(m, n, h, w, x.dtype, x.device.index)
            try:
                for k, v in self.options.items():
                    p.set_option(k, v)
            except Exception:
                p.close()                      # an option stored before any plan existed turned out to be invalid
                raise
            self.plans[key] = p
        return p

    def __call__(self, x):
        p = self.plan_for(x)
        xc = x if x.is_contiguous() else x.contiguous()
        y = torch.empty((x.shape[0], 1, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device)
        p.forward(xc, y, torch.cuda.current_stream(x.device).cuda_stream)
        return y
