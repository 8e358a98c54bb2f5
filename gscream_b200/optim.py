"""Fused Adam: a drop-in for the `torch.optim.Adam(l, lr=0.0, eps=1e-15)` GaussianModel.training_setup builds
(W-Ted/GScream scene/gaussian_model.py:374-407) and train.py:611 steps — SURVEY.md section 8f, rank 4 (second half).

`Adam` subclasses torch.optim.Optimizer and keeps torch.optim.Adam's constructor arguments, `param_groups` entries ('lr', 'betas',
'eps', 'weight_decay', plus the caller's own keys such as "name") and per-parameter state layout ('step', 'exp_avg', 'exp_avg_sq'),
so `update_learning_rate` (gaussian_model.py:460-499) and the densification code that edits the state in place
(`replace_tensor_to_optimizer`, `cat_tensors_to_optimizer`, `_prune_anchor_optimizer`, gaussian_model.py:689-781) work unchanged,
and state_dict()s are interchangeable with torch.optim.Adam's.  Only `step()` differs: one `gsr_adam_step` launch (C ABI,
include/gsr_b200.h) over every tensor of every group instead of ~10 foreach launches.  No CPU / eager fallback: CPU parameters raise.
"""
import ctypes

import torch

from . import _lib


class _AdamTensor(ctypes.Structure):
    """gsr_adam_tensor of include/gsr_b200.h."""
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p), ("exp_avg_sq", ctypes.c_void_p),
                ("numel", ctypes.c_int64), ("step", ctypes.c_int64),
                ("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double), ("eps", ctypes.c_double),
                ("weight_decay", ctypes.c_double)]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if amsgrad or maximize:
            raise NotImplementedError("fused Adam implements amsgrad=False, maximize=False (what GaussianModel.training_setup uses)")
        # the same group keys as torch.optim.Adam, so state_dict()s load either way (the implementation switches are inert here)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                                      capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False))

    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if "step" not in st:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)     # torch.optim.Adam's layout (host tensor when not capturable)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        per_device = {}
        keep = []            # contiguous gradient copies must outlive the launch
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise TypeError("fused Adam needs float32 CUDA parameters (got %s on %s); there is no CPU path" % (p.dtype, p.device))
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                if not p.is_contiguous():
                    raise ValueError("fused Adam needs contiguous parameters")
                st = self._init_state(p)
                for k in ("exp_avg", "exp_avg_sq"):
                    if not st[k].is_contiguous() or st[k].shape != p.shape or st[k].device != p.device or st[k].dtype != torch.float32:
                        raise ValueError("optimizer state %r does not match its parameter (shape %s vs %s)" % (k, tuple(st[k].shape), tuple(p.shape)))
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                per_device.setdefault(p.device, []).append(
                    _AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                                int(round(float(st["step"]))), float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                float(group["weight_decay"])))
        lib = _lib.load() if per_device else None
        for dev, rows in per_device.items():
            table = (_AdamTensor * len(rows))(*rows)
            with torch.cuda.device(dev):
                _lib.check(lib.gsr_adam_step(len(rows), ctypes.addressof(table), torch.cuda.current_stream(dev).cuda_stream))
        return loss
