"""Fused anchor -> neural-Gaussian decode: drop-in for `generate_neural_gaussians`
(gaussian_renderer/__init__.py:18-102 of W-Ted/GScream), SURVEY.md section 8f ranks 1-2.

Same signature and return tuples as the reference function, so `gaussian_renderer.render()` can import it in place
of its own; every byte of compute goes through the C ABI (`gsr_decode_stage1/2`, `gsr_decode_backward` of
include/gsr_b200.h).  Torch owns the memory, names the stream and provides the autograd node.  There is no CPU or
eager-torch fallback: unsupported configurations (use_feat_bank, feat_dim != 32, n_offsets > 16) raise.
"""
import ctypes
import itertools
import threading

import torch

from . import _lib

_PINNED = {}
_PINNED_LOCK = threading.Lock()
_PINNED_SLOTS = 64


def _pinned_counts(device):
    """Two pinned int64 (n_vis, P) for one call: slots rotate through a small per-device ring so that concurrent calls on one
    device never share counters (a call reads its own values before it returns)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _PINNED_LOCK:
        if key not in _PINNED:
            _PINNED[key] = (torch.zeros(2 * _PINNED_SLOTS, dtype=torch.int64).pin_memory(), itertools.count())
        ring, counter = _PINNED[key]
        i = 2 * (next(counter) % _PINNED_SLOTS)
        return ring[i:i + 2]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError("%s must be a float32 CUDA tensor (got %s on %s); there is no CPU decode" % (name, t.dtype, t.device))
    return t.contiguous()


def _zeros_like_many(tensors):
    """Zero-filled fp32 tensors shaped like `tensors`, as contiguous views (256-B aligned) of one flat zero-filled buffer."""
    offsets, total = [], 0
    for t in tensors:
        offsets.append(total)
        total += (t.numel() + 63) // 64 * 64
    flat = torch.zeros(max(total, 1), dtype=torch.float32, device=tensors[0].device)
    return [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offsets, tensors)]


class _DecodeAnchors(torch.autograd.Function):
    """inputs: anchor[A,3] feat[A,32] offset[A,k,3] scaling[A,6] campos[3] visible_mask (bool[A] or None) + the 16 MLP
    tensors ({opacity, uncertainty, cov, colour} x {w1, b1, w2, b2}).
    outputs: xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask — gaussian_renderer/__init__.py:96-98."""

    @staticmethod
    def forward(ctx, anchor, feat, offset, scaling, campos, visible_mask, *mlp):
        lib = _lib.load()
        if len(mlp) != 16:
            raise ValueError("expected 16 MLP tensors, got %d" % len(mlp))
        anchor, feat, offset, scaling = _f32c(anchor, "anchor"), _f32c(feat, "anchor_feat"), _f32c(offset, "offset"), _f32c(scaling, "scaling")
        campos = _f32c(campos, "camera_center").reshape(-1)
        mlp = tuple(_f32c(t, "mlp parameter") for t in mlp)
        A, feat_dim = feat.shape
        k = offset.shape[1]
        if not lib.gsr_decode_supported(int(feat_dim), int(k)):
            raise NotImplementedError("fused decode supports feat_dim == 32 and 1 <= n_offsets <= 16 (got %d, %d)" % (feat_dim, k))
        if anchor.shape != (A, 3) or offset.shape != (A, k, 3) or scaling.shape != (A, 6):
            raise ValueError("anchor / offset / scaling shapes do not match anchor_feat")
        n_out = (k, k, 7 * k, 3 * k)
        for m in range(4):
            w1, b1, w2, b2 = mlp[4 * m:4 * m + 4]
            if w1.shape != (32, 36) or b1.shape != (32,) or w2.shape != (n_out[m], 32) or b2.shape != (n_out[m],):
                raise ValueError("MLP %d does not have the 36 -> 32 -> %d shape of scene/gaussian_model.py:118-144" % (m, n_out[m]))
        dev = feat.device
        vm = None
        if visible_mask is not None:
            vm = visible_mask.to(torch.bool).contiguous()
            if vm.shape != (A,):
                raise ValueError("visible_mask must have shape [A]")
        scratch = torch.empty(max(int(lib.gsr_decode_scratch_bytes(A)), 16), dtype=torch.uint8, device=dev)
        nop_full = torch.empty(A * k, dtype=torch.float32, device=dev)
        mask_full = torch.empty(A * k, dtype=torch.bool, device=dev)
        counts = _pinned_counts(dev)
        params = _ptr_array(mlp)
        with torch.cuda.device(dev):
            stream = _stream()
            _lib.check(lib.gsr_decode_stage1(A, feat_dim, k, anchor.data_ptr(), feat.data_ptr(), vm.data_ptr() if vm is not None else None,
                                             campos.data_ptr(), params, scratch.data_ptr(), scratch.numel(), nop_full.data_ptr(),
                                             mask_full.data_ptr(), counts.data_ptr(), stream))
            # Output sizes are data dependent (the reference syncs on every boolean index).  The counts travel to a pinned
            # buffer behind stage 1; stage 2 is launched right away into outputs with the capacity of all A * k offsets (with a
            # visibility mask it reads n_vis from device memory), and the host waits on an EVENT behind the counts only — the
            # GPU keeps running stage 2 — before narrowing the outputs to their first P rows.
            counted = torch.cuda.Event()
            counted.record()
            cap = A * k
            xyz = torch.empty(cap, 3, dtype=torch.float32, device=dev)
            color = torch.empty(cap, 3, dtype=torch.float32, device=dev)
            opacity = torch.empty(cap, 1, dtype=torch.float32, device=dev)
            uncertainty = torch.empty(cap, 1, dtype=torch.float32, device=dev)
            out_scaling = torch.empty(cap, 3, dtype=torch.float32, device=dev)
            rot = torch.empty(cap, 4, dtype=torch.float32, device=dev)
            if cap > 0:
                _lib.check(lib.gsr_decode_stage2(A, feat_dim, k, -1 if vm is not None else A, cap, anchor.data_ptr(), feat.data_ptr(),
                                                 offset.data_ptr(), scaling.data_ptr(), campos.data_ptr(), params, scratch.data_ptr(),
                                                 scratch.numel(), nop_full.data_ptr(), xyz.data_ptr(), color.data_ptr(), opacity.data_ptr(),
                                                 uncertainty.data_ptr(), out_scaling.data_ptr(), rot.data_ptr(), stream))
            counted.synchronize()
            n_vis, P = int(counts[0]), int(counts[1])
            xyz, color, opacity, uncertainty, out_scaling, rot = (t[:P] for t in (xyz, color, opacity, uncertainty, out_scaling, rot))
        neural_opacity = nop_full[:n_vis * k].view(-1, 1)
        mask = mask_full[:n_vis * k]
        ctx.save_for_backward(anchor, feat, offset, scaling, campos, scratch, *mlp)
        ctx.dims = (A, feat_dim, k, n_vis, P)
        ctx.mark_non_differentiable(mask)
        return xyz, color, opacity, uncertainty, out_scaling, rot, neural_opacity, mask

    @staticmethod
    def backward(ctx, d_xyz, d_color, d_opacity, d_uncertainty, d_scaling, d_rot, d_nop, _d_mask):
        lib = _lib.load()
        anchor, feat, offset, scaling, campos, scratch = ctx.saved_tensors[:6]
        mlp = ctx.saved_tensors[6:]
        A, feat_dim, k, n_vis, P = ctx.dims
        dev = feat.device

        def up(g):
            return None if g is None else g.to(torch.float32).contiguous()

        ups = [up(g) for g in (d_xyz, d_color, d_opacity, d_uncertainty, d_scaling, d_rot, d_nop)]
        # twenty zero-filled outputs carved out of ONE zero-filled allocation (one fill launch instead of twenty)
        g_anchor, g_feat, g_offset, g_scaling, *g_mlp = _zeros_like_many((anchor, feat, offset, scaling) + tuple(mlp))
        with torch.cuda.device(dev):
            _lib.check(lib.gsr_decode_backward(A, feat_dim, k, n_vis, P, anchor.data_ptr(), feat.data_ptr(), offset.data_ptr(),
                                               scaling.data_ptr(), campos.data_ptr(), _ptr_array(mlp), scratch.data_ptr(), scratch.numel(),
                                               *[g.data_ptr() if g is not None and g.numel() else None for g in ups],
                                               g_anchor.data_ptr(), g_feat.data_ptr(), g_offset.data_ptr(), g_scaling.data_ptr(),
                                               _ptr_array(g_mlp), _stream()))
        return (g_anchor, g_feat, g_offset, g_scaling, None, None) + tuple(g_mlp)


def _mlp_params(seq):
    """(w1, b1, w2, b2) of an nn.Sequential(Linear, ReLU, Linear[, activation]) — scene/gaussian_model.py:118-144."""
    return seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias


def decode_anchors(anchor, anchor_feat, offset, scaling, camera_center, visible_mask, mlp_opacity, mlp_uncertainty, mlp_cov, mlp_color):
    """Tensor-level entry point: returns (xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask)."""
    params = _mlp_params(mlp_opacity) + _mlp_params(mlp_uncertainty) + _mlp_params(mlp_cov) + _mlp_params(mlp_color)
    return _DecodeAnchors.apply(anchor, anchor_feat, offset, scaling, camera_center, visible_mask, *params)


def generate_neural_gaussians(viewpoint_camera, pc, visible_mask=None, is_training=False):
    """Drop-in for gaussian_renderer/__init__.py:18-102 (same arguments, same return tuples)."""
    if getattr(pc, "use_feat_bank", False):
        raise NotImplementedError("fused decode: use_feat_bank=True (gaussian_renderer/__init__.py:39-49) is not supported")
    out = decode_anchors(pc.get_anchor, pc._anchor_feat, pc._offset, pc.get_scaling, viewpoint_camera.camera_center, visible_mask,
                         pc.get_opacity_mlp, pc.get_uncertainty_mlp, pc.get_cov_mlp, pc.get_color_mlp)
    xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask = out
    if is_training:
        return xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask
    return xyz, color, opacity, uncertainty, scaling, rot
