"""Fused densification statistics: drop-in for `GaussianModel.training_statis` (scene/gaussian_model.py:729-757 of
W-Ted/GScream), SURVEY.md section 8f rank 4 (first half).

Same arguments as the reference method (the model is the first argument, so it can be bound in its place:
`GaussianModel.training_statis = gscream_b200.stats.training_statis`); updates `opacity_accum`, `anchor_demon`,
`offset_gradient_accum` and `offset_denom` in place through `gsr_training_statis` (two scans + one kernel instead of ~15
boolean-mask index kernels, each of which synchronises the host).  No CPU fallback.
"""
import torch

from . import _lib


def _u8(t, name):
    if not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor; there is no CPU path" % name)
    t = t.contiguous()
    return t.view(torch.uint8) if t.dtype == torch.bool else t.to(torch.uint8)


def training_statis(self, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
    lib = _lib.load()
    k = int(self.n_offsets)
    acc = self.opacity_accum
    if not acc.is_cuda or acc.dtype != torch.float32:
        raise TypeError("the statistics buffers must be float32 CUDA tensors; there is no CPU path")
    for name in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        if not getattr(self, name).is_contiguous():
            raise ValueError("%s must be contiguous (it is updated in place)" % name)
    A = int(acc.shape[0])
    grad = viewspace_point_tensor.grad
    if grad is None:
        raise RuntimeError("viewspace_point_tensor has no .grad (call after loss.backward(), with retain_grad)")
    grad = grad.detach().to(torch.float32).contiguous()
    nop = opacity.detach().to(torch.float32).contiguous().view(-1)
    n_vis, P = nop.numel() // k, int(grad.shape[0])
    vis, sel, upd = _u8(anchor_visible_mask, "anchor_visible_mask"), _u8(offset_selection_mask, "offset_selection_mask"), _u8(update_filter, "update_filter")
    if vis.numel() != A or sel.numel() != n_vis * k or upd.numel() != P or self.offset_denom.numel() != A * k:
        raise ValueError("mask / buffer sizes are inconsistent")
    dev = acc.device
    scratch = torch.empty(max(int(lib.gsr_training_statis_scratch_bytes(A, k)), 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.gsr_training_statis(A, k, n_vis, P, vis.data_ptr(), sel.data_ptr(), upd.data_ptr() if P else None, nop.data_ptr(),
                                           grad.data_ptr() if P else None, self.opacity_accum.data_ptr(), self.anchor_demon.data_ptr(),
                                           self.offset_gradient_accum.data_ptr(), self.offset_denom.data_ptr(), scratch.data_ptr(),
                                           scratch.numel(), torch.cuda.current_stream().cuda_stream))
