"""Fused Adam (gscream_b200/optim.py + gsr_optim.cu) against torch.optim.Adam — the class the
reference instantiates (scene/gaussian_model.py:407, `torch.optim.Adam(l, lr=0.0, eps=1e-15)`) and steps (train.py:611).

torch is a third-party dependency of the reference (gscream.yaml pins torch 1.12.1; the image has 2.11): its Adam IS the oracle
here, run in fp64 on the CPU.  CPU tests cover the host logic (group keys, state layout, densification-style state surgery, the
descriptor struct, loud failure without CUDA) and pin a numpy restatement of the kernel's order of operations to torch's fp32
result; GPU tests compare the CUDA step with torch.optim.Adam (fp32 on the GPU and fp64 on the CPU) through the C ABI.

Tolerance: |ours - adam64| <= 2 |torch32 - adam64| + 1e-6 max|adam64|  per tensor, parameters and both moments.
"""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _groups(tensors, dtype, device):
    """Parameter groups shaped like training_setup's: named groups, one lr each, lr=0.0 as the default."""
    ps = [torch.nn.Parameter(t.to(dtype=dtype, device=device).clone()) for t in tensors]
    groups = [{"params": [ps[0]], "lr": 1.6e-3, "name": "anchor"}, {"params": [ps[1]], "lr": 1e-2, "name": "offset"},
              {"params": [ps[2]], "lr": 7.5e-3, "name": "anchor_feat"}, {"params": ps[3:5], "lr": 2e-3, "name": "mlp_opacity"},
              {"params": ps[5:], "lr": 4e-3, "name": "rest", "weight_decay": 0.0}]
    return ps, groups


def _tensors(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1000, 3), (1000, 10, 3), (1000, 32), (32, 35), (32,), (1,), (4097,), (3, 5, 7), (10, 32), (8193, 2)]
    return [torch.randn(*s, generator=g) for s in shapes], shapes


def _grads(shapes, step, seed=0):
    g = torch.Generator().manual_seed(1000 + 17 * seed + step)
    out = [torch.randn(*s, generator=g) * (10.0 ** ((i % 5) - 3)) for i, s in enumerate(shapes)]
    out[0][::7] = 0.0            # rows of invisible anchors receive exactly zero gradients
    return out


def _run(opt_cls, dtype, device, steps=6, skip_grad_of=(8,), wd_group=None, **kw):
    tensors, shapes = _tensors()
    ps, groups = _groups(tensors, dtype, device)
    if wd_group is not None:
        groups[wd_group]["weight_decay"] = 0.01
    opt = opt_cls(groups, lr=0.0, eps=1e-15, **kw)
    for it in range(steps):
        for i, (p, g) in enumerate(zip(ps, _grads(shapes, it))):
            p.grad = None if (i in skip_grad_of and it % 2 == 0) else g.to(dtype=dtype, device=device)
        for group in opt.param_groups:     # update_learning_rate (gaussian_model.py:460-499) rewrites 'lr' every iteration
            if group["name"] in ("anchor", "offset"):
                group["lr"] = group["lr"] * 0.97
        opt.step()
        opt.zero_grad(set_to_none=True)
    return ps, opt


def test_group_keys_and_state_layout_match_torch_adam():
    from gscream_b200 import optim
    w = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5))]
    mk = lambda cls: cls([{"params": [w[0]], "lr": 0.1, "name": "anchor"}, {"params": [w[1]], "name": "mlp"}], lr=0.0, eps=1e-15)
    ours, ref = mk(optim.Adam), mk(torch.optim.Adam)
    assert set(ours.param_groups[0].keys()) == set(ref.param_groups[0].keys())
    for a, b in zip(ours.param_groups, ref.param_groups):
        for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "name"):
            assert a[k] == b[k], k
    # state dicts load either way
    for p in w:
        p.grad = torch.ones_like(p)
    ref.step()
    ours.load_state_dict(ref.state_dict())
    st = ours.state[w[0]]
    assert set(st.keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 1.0
    ref.load_state_dict(ours.state_dict())
    with pytest.raises(NotImplementedError):
        optim.Adam(w, amsgrad=True)
    with pytest.raises(ValueError):
        optim.Adam(w, lr=-1.0)


def test_fused_adam_fails_loudly_on_cpu_parameters():
    from gscream_b200 import optim
    w = torch.nn.Parameter(torch.zeros(3))
    opt = optim.Adam([w], lr=0.1)
    opt.step()                    # nothing has a gradient: a no-op like torch's, no library needed
    w.grad = torch.ones(3)
    with pytest.raises(TypeError):
        opt.step()
    assert float(w.detach().abs().sum()) == 0.0


def test_descriptor_struct_matches_header():
    from gscream_b200 import optim
    hdr = open(os.path.join(HERE, "..", "include", "gsr_b200.h")).read()
    body = re.search(r"typedef struct gsr_adam_tensor \{(.*?)\} gsr_adam_tensor;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip(" *") for decl in body.split(";") if decl.strip() for n in decl.split(",")]
    names = [n.split()[-1].lstrip("*") for n in names]
    assert names == [f[0] for f in optim._AdamTensor._fields_]
    assert ctypes.sizeof(optim._AdamTensor) == 88 and optim._AdamTensor.numel.offset == 32 and optim._AdamTensor.lr.offset == 48


def _kernel_order_numpy(p, g, m, v, lr, beta1, beta2, eps, wd, step):
    """fp32 restatement of adam_update() in gsr_optim.cu with the host-side bias corrections of gsr_adam_step()."""
    f = np.float32
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    step_size, bc2s = f(lr / bc1), f(np.sqrt(bc2))
    if wd != 0:
        g = f(wd) * p + g
    m = m + (g - m) * f(1.0 - beta1)
    v = v * f(beta2) + f(1.0 - beta2) * g * g
    denom = np.sqrt(v) / bc2s + f(eps)
    p = p - step_size * (m / denom)
    return p, m, v


def test_kernel_order_of_operations_matches_torch_adam_fp32_and_fp64():
    tensors, shapes = _tensors(3)
    p0 = tensors[2]
    for wd in (0.0, 0.01):
        ref32 = torch.nn.Parameter(p0.clone())
        ref64 = torch.nn.Parameter(p0.double().clone())
        o32 = torch.optim.Adam([ref32], lr=3e-3, eps=1e-15, weight_decay=wd)
        o64 = torch.optim.Adam([ref64], lr=3e-3, eps=1e-15, weight_decay=wd)
        p, m, v = p0.numpy().copy(), np.zeros_like(p0.numpy()), np.zeros_like(p0.numpy())
        for it in range(1, 8):
            g = _grads(shapes, it, 3)[2]
            ref32.grad, ref64.grad = g.clone(), g.double()
            o32.step(); o64.step()
            p, m, v = _kernel_order_numpy(p, g.numpy(), m, v, 3e-3, 0.9, 0.999, 1e-15, wd, it)
        r64 = ref64.detach().numpy()
        spread = np.abs(ref32.detach().numpy() - r64).max()
        assert np.abs(p - r64).max() <= 2 * spread + 1e-6 * np.abs(r64).max()
        m64, v64 = o64.state[ref64]["exp_avg"].numpy(), o64.state[ref64]["exp_avg_sq"].numpy()
        np.testing.assert_allclose(m, m64, rtol=0, atol=1e-6 * np.abs(m64).max())     # fp32 moments: absolute, relative to the tensor's scale
        np.testing.assert_allclose(v, v64, rtol=0, atol=1e-6 * np.abs(v64).max())


def _compare(ours_ps, ours_opt, t32_ps, t32_opt, r64_ps, r64_opt):
    for i, (a, b, c) in enumerate(zip(ours_ps, t32_ps, r64_ps)):
        for name, x, y, z in (("param", a.detach(), b.detach(), c.detach()),
                              ("exp_avg", ours_opt.state[a].get("exp_avg"), t32_opt.state[b].get("exp_avg"), r64_opt.state[c].get("exp_avg")),
                              ("exp_avg_sq", ours_opt.state[a].get("exp_avg_sq"), t32_opt.state[b].get("exp_avg_sq"), r64_opt.state[c].get("exp_avg_sq"))):
            x, y, z = x.double().cpu().numpy(), y.double().cpu().numpy(), z.double().cpu().numpy()
            tol = 2.0 * np.abs(y - z).max() + 1e-6 * max(np.abs(z).max(), 1e-30)
            assert np.abs(x - z).max() <= tol, "tensor %d %s: err %.3e tol %.3e" % (i, name, np.abs(x - z).max(), tol)
        assert float(ours_opt.state[a]["step"]) == float(r64_opt.state[c]["step"])


@pytest.mark.gpu
@pytest.mark.parametrize("wd_group", (None, 4))
def test_gpu_fused_adam_matches_torch_adam(wd_group):
    from gscream_b200 import _lib, optim
    lib = _lib.load()
    n0 = lib.gsr_launch_count(0)
    ours = _run(optim.Adam, torch.float32, "cuda", wd_group=wd_group)
    assert lib.gsr_launch_count(0) - n0 == 6, "one launch per step for 10 tensors"
    t32 = _run(torch.optim.Adam, torch.float32, "cuda", wd_group=wd_group)
    r64 = _run(torch.optim.Adam, torch.float64, "cpu", wd_group=wd_group)
    _compare(*ours, *t32, *r64)


@pytest.mark.gpu
def test_gpu_fused_adam_many_tensors_unaligned_and_state_surgery():
    """> 24 tensors (two launches), parameters at 4-byte-aligned offsets (scalar path), and the densification code's state
    surgery (cat_tensors_to_optimizer, gaussian_model.py:705-727): the moments are extended with zeros, `step` is kept."""
    from gscream_b200 import optim
    g = torch.Generator().manual_seed(5)
    base = [torch.randn(37 + 13 * i, generator=g) for i in range(30)]

    def build(cls, dtype, device):
        ps = []
        for i, t in enumerate(base):
            store = torch.zeros(t.numel() + 3, dtype=dtype, device=device)
            view = store[1 + (i % 3):1 + (i % 3) + t.numel()]          # storage offsets 1, 2, 3 elements: not 16-byte aligned
            view.copy_(t.to(dtype=dtype, device=device))
            ps.append(torch.nn.Parameter(view))
        return ps, cls([{"params": [p], "lr": 1e-3 * (1 + i % 4), "name": "t%d" % i} for i, p in enumerate(ps)], lr=0.0, eps=1e-15)

    def surgery(ps, opt, dtype, device):
        group = opt.param_groups[0]
        old = group["params"][0]
        st = opt.state.get(old)
        ext = torch.full((11,), 0.5, dtype=dtype, device=device)
        st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)))
        st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)))
        del opt.state[old]
        new = torch.nn.Parameter(torch.cat((old.detach(), ext)))
        group["params"][0] = new
        opt.state[new] = st
        ps[0] = new

    runs = []
    for cls, dtype, device in ((optim.Adam, torch.float32, "cuda"), (torch.optim.Adam, torch.float32, "cuda"), (torch.optim.Adam, torch.float64, "cpu")):
        ps, opt = build(cls, dtype, device)
        for it in range(4):
            if it == 2:
                surgery(ps, opt, dtype, device)
            gg = torch.Generator().manual_seed(100 + it)
            for p in ps:
                p.grad = torch.randn(p.numel(), generator=gg).to(dtype=dtype, device=device)
            opt.step()
        runs += [ps, opt]
    _compare(*runs)
