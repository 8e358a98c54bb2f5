#!/usr/bin/env python
"""bench.py — forward+backward views/s of the Gaussian-rasterizer hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config2|config4]

One "step" = one view per GPU: forward (cull/project, bin, sort, blend) + backward (blend backward, per-Gaussian
backward) of the named synthetic workload, gradients accumulated into one flat bucket, plus — at N > 1 — the single
NCCL sum-allreduce of that bucket.  Default workload: BASELINE.json configs[2] ("config3": 1M Gaussians, 1920x1080,
32 feature channels + depth + uncertainty), the configuration the metric is quoted on; it fits one GPU.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline` is for the dominant kernel (blend
backward), `cpu_baseline` is the CPU oracle port timed on a bounded sample, `e2e` goes through the public
GaussianRasterizer API with every input coming from pinned HOST memory inside the timed region.

--impl reference times the reference's own CUDA build (oracle/_ref, compiled from /root/reference by
oracle/build_ref.py; the reference ships no CPU rasterizer, so per BASELINE.json's north_star that build on the same
GPU — with the host core count recorded — is the baseline arm) on the same config through its own Python API.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the `ncu --set full` capture recorded in
    profiles/traffic.json — valid only for the kernel build it was captured from (keyed by the digest of the blend kernels'
    sources + compiler flags, gscream_b200/_build.blend_digest); any other build reports null rather than a stale constant."""
    try:
        from gscream_b200 import _build
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            table = json.load(fh)
        entry = table.get(_build.blend_digest(), {}).get(kernel)
        return (int(entry["dram_bytes"]), entry.get("source")) if entry else (None, None)
    except Exception:
        return None, None


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _timed(fn_step, steps, warmup, world, sampler=None):
    """W untimed steps, then exactly K steps bracketed by barrier + synchronize; device time, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clocks


def algorithmic_bytes_blend_backward(C, W, H, R, V):
    """B1 + 4 (K+6) V  with  B1 = 8 Tn + 28 R + 4 K R + 4 (K+2) N."""
    K = C + 2
    Tn = ((W + 15) // 16) * ((H + 15) // 16)
    N = W * H
    return 8 * Tn + 28 * R + 4 * K * R + 4 * (K + 2) * N + 4 * (K + 6) * V


def cpu_baseline_port(cfg):
    """CPU oracle (plain C, 1 thread) on a bounded sample of the workload: 1/4 of the Gaussians at 1/2 x 1/2 of the
    image (same splat density and per-pixel depth complexity, ~10 s of CPU work); reported as full-workload-equivalent views/s."""
    from oracle.oracle import Oracle
    from gscream_b200 import scenes
    P, W, H, C = cfg["P"] // 4, cfg["W"] // 2, cfg["H"] // 2, cfg["C"]
    s = scenes.make_scene(P, W, H, C, cfg["seed"])
    cam = scenes.make_camera(W, H)
    g = scenes.make_upstream_grads(C, W, H, cfg["seed"])
    o = Oracle("f32")
    a = dict(means3D=s["means3D"].numpy(), colors_precomp=s["colors"].numpy(), opacities=s["opacities"].numpy(),
             uncertainties=s["uncertainties"].numpy(), scales=s["scales"].numpy(), rotations=s["rotations"].numpy(),
             viewmatrix=cam["viewmatrix"].numpy(), projmatrix=cam["projmatrix"].numpy(), bg=s["bg"].numpy(), W=W, H=H,
             tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"])
    t0 = time.perf_counter()
    f = o.forward(**a)
    o.backward(f, means3D=a["means3D"], colors_precomp=a["colors_precomp"], scales=a["scales"], rotations=a["rotations"],
               viewmatrix=a["viewmatrix"], projmatrix=a["projmatrix"], bg=a["bg"], W=W, H=H, tanfovx=a["tanfovx"], tanfovy=a["tanfovy"],
               dL_dcolor=g[0].numpy(), dL_ddepth=g[1].numpy(), dL_dunc=g[2].numpy())
    dt = time.perf_counter() - t0
    return {"value": (1.0 / dt) / 4.0, "unit": "views/s", "cores": 1, "kind": "port",
            "sample": "1/4 of the workload: %d Gaussians at %dx%d (same density), one fwd+bwd view in %.2f s on 1 thread; value = measured/4" % (P, W, H, dt)}


def run_ours(args, cfg, rank, world, local):
    import torch.distributed as dist
    from gscream_b200 import _lib, scenes
    from gscream_b200 import rasterizer as ours
    from gscream_b200.dist import GradBucket, allreduce_bucket, render_views_into_bucket
    lib = _lib.load()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    P, W, H, C, seed = cfg["P"], cfg["W"], cfg["H"], cfg["C"], cfg["seed"]
    scene_cpu = scenes.make_scene(P, W, H, C, seed, scale_mult=args.scale_mult)  # identical bits on every rank
    yaw = (rank - (world - 1) / 2.0) * 3.0                                # every rank renders its own view
    if args.yaw_deg is not None:
        yaw = args.yaw_deg
    cam_cpu = scenes.make_camera(W, H, yaw_deg=yaw)
    grads_cpu = scenes.make_upstream_grads(C, W, H, seed + rank)
    scene = {k: v.to(dev) for k, v in scene_cpu.items()}
    cam = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in cam_cpu.items()}
    ups = [tuple(g.to(dev) for g in grads_cpu)]
    bucket = GradBucket(P, C, device=dev)
    info = {}

    def step():
        # (no bucket.zero_(): the first view's backward overwrites the bucket, gscream_b200/dist.py)
        outs = render_views_into_bucket(scene, [cam], ups, bucket, keep_outputs=True, overwrite=True)
        info["R"] = outs[0][4]
        info["radii"] = outs[0][3]
        allreduce_bucket(bucket)

    # ---- device-resident throughput (`value`) + per-stage device times ----
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    lib.gsr_profile_enable(1)
    lib.gsr_launch_count(1)
    ms, clocks = _timed(step, args.steps, 0, world, ClockSampler(local) if rank == 0 else None)
    launches = int(lib.gsr_launch_count(0))
    stage_ms = {}
    buf = np.zeros(256, np.float32)
    for sid, name in enumerate(("preprocess", "depth_order_scan", "binning", "blend_forward", "blend_backward", "gaussian_backward")):
        n = lib.gsr_profile_read(sid, buf.ctypes.data, 256)
        stage_ms[name] = float(buf[:n].mean()) if n > 0 else None
    lib.gsr_profile_enable(0)
    stage_ms_rank0 = dict(stage_ms)
    if world > 1:   # ms_per_step is the max over ranks; so are the stage times (rank 0's own are kept beside them)
        names = sorted(stage_ms)
        t = torch.tensor([stage_ms[k] or 0.0 for k in names], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage_ms = {k: float(v) for k, v in zip(names, t.tolist())}
    ms_per_step = ms / args.steps
    value = world * 1000.0 / ms_per_step
    R = int(info["R"])
    V = int((info["radii"] > 0).sum().item())

    if args.value_only:
        if rank == 0:
            return {"metric": "fwd+bwd views/sec", "value": value, "unit": "views/s", "n_gpus": world, "ms_per_step": ms_per_step, "num_rendered": R, "visible": V,
                    "stage_ms": stage_ms, "stage_ms_rank0": stage_ms_rank0, "note": "--value-only development run"}
        return None
    # ---- end to end through the public API, inputs from pinned host memory every step ----
    keys = ("means3D", "colors", "opacities", "uncertainties", "scales", "rotations")
    host = {k: scene_cpu[k].pin_memory() for k in keys}
    host_cam = {k: cam_cpu[k].pin_memory() for k in ("viewmatrix", "projmatrix", "campos")}
    host_grads = [g.pin_memory() for g in grads_cpu]
    h2d_bytes = sum(t.numel() * 4 for t in host.values()) + sum(t.numel() * 4 for t in host_cam.values()) + sum(g.numel() * 4 for g in host_grads)
    loss_host = torch.zeros(1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [dict() for _ in range(2)]                                      # double-buffered device staging
    for sl in slots:
        for k in keys:
            sl[k] = torch.empty_like(scene[k])
        for k in host_cam:
            sl[k] = torch.empty_like(cam[k])
        sl["g"] = [torch.empty_like(g) for g in ups[0]]
        sl["ready"] = torch.cuda.Event()
        sl["free"] = torch.cuda.Event()
        sl["free"].record()
    state = {"i": 0, "primed": False}

    def upload(sl):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])
            for k in keys:
                sl[k].copy_(host[k], non_blocking=True)
            for k in host_cam:
                sl[k].copy_(host_cam[k], non_blocking=True)
            for d, h in zip(sl["g"], host_grads):
                d.copy_(h, non_blocking=True)
            sl["ready"].record(copy_stream)

    def e2e_step():
        # step i consumes slot i%2 (uploaded during step i-1) and starts the upload of step i+1's inputs
        if not state["primed"]:
            upload(slots[0])
            state["primed"] = True
        sl = slots[state["i"] % 2]
        upload(slots[(state["i"] + 1) % 2])
        cur = torch.cuda.current_stream()
        cur.wait_event(sl["ready"])
        leaves = {k: sl[k].requires_grad_(True) for k in keys}
        m2d = torch.zeros_like(sl["means3D"], requires_grad=True)
        st = ours.GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], scene["bg"], 1.0, sl["viewmatrix"], sl["projmatrix"],
                                                1, sl["campos"], False, False)
        color, depth, unc, radii = ours.GaussianRasterizer(st)(
            means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], uncertainties=leaves["uncertainties"],
            colors_precomp=leaves["colors"], scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward((color, depth, unc), tuple(sl["g"]))
        if world > 1:
            for k in keys:
                dist.all_reduce(leaves[k].grad, op=dist.ReduceOp.SUM)
        loss = (color.detach() * sl["g"][0]).sum() + leaves["means3D"].grad.abs().sum()
        loss_host.copy_(loss.reshape(1), non_blocking=True)                 # the step's result goes back to the host
        for k in keys:
            sl[k].requires_grad_(False)
            sl[k].grad = None
        sl["free"].record(cur)
        state["i"] += 1

    e2e_all_ms, _ = _timed(e2e_step, args.steps, max(3, args.warmup), world)
    torch.cuda.synchronize()
    e2e_all_value = world * 1000.0 * args.steps / e2e_all_ms
    del slots

    # ---- end to end, training-view forms: public API forward + autograd backward, the Gaussian arrays resident like model
    # weights, the camera copied from pinned host memory every step, a result scalar read back.  Two variants of where the
    # view's supervision (the upstream-gradient planes: C colour + depth + uncertainty) lives:
    #   resident  — on the device, as in GScream, whose cameras keep every view's images on the GPU (scene/cameras.py): `e2e`;
    #   from host — copied from pinned host memory every step too (282 MB at config3: a measurement of the PCIe link). ----
    def view_form(supervision_from_host):
        nslots = 2
        vs = [dict() for _ in range(nslots)]
        for sl in vs:
            for k in host_cam:
                sl[k] = torch.empty_like(cam[k])
            sl["g"] = [torch.empty_like(g) for g in ups[0]] if supervision_from_host else list(ups[0])
            sl["ready"] = torch.cuda.Event()
            sl["free"] = torch.cuda.Event()
            sl["free"].record()
        res = {k: scene[k].clone().requires_grad_(True) for k in keys}
        m2d_res = torch.zeros_like(res["means3D"], requires_grad=True)
        st8 = {"i": 0, "primed": False}

        def upload_view(sl):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(sl["free"])
                for k in host_cam:
                    sl[k].copy_(host_cam[k], non_blocking=True)
                if supervision_from_host:
                    for d, h in zip(sl["g"], host_grads):
                        d.copy_(h, non_blocking=True)
                sl["ready"].record(copy_stream)

        def step_fn():
            if not st8["primed"]:
                upload_view(vs[0])
                st8["primed"] = True
            sl = vs[st8["i"] % 2]
            upload_view(vs[(st8["i"] + 1) % 2])
            cur = torch.cuda.current_stream()
            cur.wait_event(sl["ready"])
            st = ours.GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], scene["bg"], 1.0, sl["viewmatrix"], sl["projmatrix"],
                                                    1, sl["campos"], False, False)
            color, depth, unc, radii = ours.GaussianRasterizer(st)(
                means3D=res["means3D"], means2D=m2d_res, opacities=res["opacities"], uncertainties=res["uncertainties"],
                colors_precomp=res["colors"], scales=res["scales"], rotations=res["rotations"])
            torch.autograd.backward((color, depth, unc), tuple(sl["g"]))
            if world > 1:   # the step's collective, through the public API: the parameter gradients of all ranks are summed
                for k in keys:
                    dist.all_reduce(res[k].grad, op=dist.ReduceOp.SUM)
            loss = (color.detach() * sl["g"][0]).sum() + res["means3D"].grad.abs().sum()
            loss_host.copy_(loss.reshape(1), non_blocking=True)
            for t in list(res.values()) + [m2d_res]:
                t.grad = None
            sl["free"].record(cur)
            st8["i"] += 1

        ms_, _ = _timed(step_fn, args.steps, max(3, args.warmup), world)
        torch.cuda.synchronize()
        return world * 1000.0 * args.steps / ms_

    cam_bytes = sum(t.numel() * 4 for t in host_cam.values())
    h2d_view_bytes = cam_bytes + sum(g.numel() * 4 for g in host_grads)
    e2e_value = view_form(False)
    e2e_sup_value = view_form(True)

    # ---- public API + autograd with everything resident on the device: the like-for-like arm against `--impl reference`
    # (which is timed exactly this way) ----
    res = {k: scene[k].clone().requires_grad_(True) for k in keys}
    m2d_res = torch.zeros_like(res["means3D"], requires_grad=True)
    st_res = ours.GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], scene["bg"], 1.0, cam["viewmatrix"], cam["projmatrix"],
                                                1, cam["campos"], False, False)
    rast_res = ours.GaussianRasterizer(st_res)

    def api_step():
        color, depth, unc, radii = rast_res(means3D=res["means3D"], means2D=m2d_res, opacities=res["opacities"], uncertainties=res["uncertainties"],
                                            shs=None, colors_precomp=res["colors"], scales=res["scales"], rotations=res["rotations"], cov3D_precomp=None)
        torch.autograd.backward((color, depth, unc), ups[0])
        if world > 1:
            for k in keys:
                dist.all_reduce(res[k].grad, op=dist.ReduceOp.SUM)
        for t in list(res.values()) + [m2d_res]:
            t.grad = None

    api_ms, _ = _timed(api_step, args.steps, max(3, args.warmup), world)
    torch.cuda.synchronize()
    api_value = world * 1000.0 * args.steps / api_ms
    del res, m2d_res

    # per-tile list length statistics (outside any timed region)
    from gscream_b200 import _C as gC
    outs = render_views_into_bucket(scene, [cam], ups, bucket.zero_(), keep_outputs=False)
    R_, col_, dep_, unc_, rad_, geom_, binning_, img_ = gC.rasterize_gaussians(
        scene["bg"], scene["means3D"], scene["colors"], scene["opacities"], scene["uncertainties"], scene["scales"], scene["rotations"], 1.0,
        torch.empty(0), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], H, W, torch.empty(0), 1, cam["campos"], False, False)
    rng = gC.debug_export(P, R_, W, H, geom_, binning_, img_)["ranges"].cpu().numpy().astype(np.int64)
    lens = rng[:, 1] - rng[:, 0]
    tile_stats = {"mean": float(lens.mean()), "p50": float(np.percentile(lens, 50)), "p90": float(np.percentile(lens, 90)),
                  "p99": float(np.percentile(lens, 99)), "max": int(lens.max())}
    del R_, col_, dep_, unc_, rad_, geom_, binning_, img_, outs

    out = None
    if rank == 0:
        peak, peak_src = _peaks()
        bytes_bwd = algorithmic_bytes_blend_backward(C, W, H, R, V)
        t_bwd = stage_ms_rank0["blend_backward"]
        bwd_kernel = "gsr::blend_backward_kernel<%d>" % C
        traffic, traffic_src = _measured_traffic(bwd_kernel + ":" + args.workload) if args.scale_mult == 1.0 else (None, None)
        achieved = bytes_bwd / (t_bwd * 1e-3) / 1e9 if t_bwd else None
        out = {
            "metric": "fwd+bwd views/sec", "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded Gaussian cloud, SURVEY.md 8d; one camera view per GPU)",
            "config": {"workload": "%s%s: %d Gaussians, %dx%d, %d feature channels + depth + uncertainty, fwd+bwd" % (
                           args.workload, "" if args.scale_mult == 1.0 else " heavy (splat scale x%g)" % args.scale_mult, P, W, H, C),
                       "views_per_step": world, "num_rendered": R, "visible": V, "instances_per_gaussian": R / max(P, 1),
                       "tile_list_len": tile_stats, "parallelism": "view-parallel dp%d" % world,
                       "l2": "working set (features 128 MB + records 64 MB + planes 282 MB x2) exceeds the 126 MB L2; no explicit flush",
                       "collective": ("1 NCCL sum-allreduce of the %.0f MB parameter-gradient bucket per step (the per-view means2D block stays local)" % (bucket.reduced_nbytes() / 1e6)) if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": cam_bytes, "d2h_bytes_per_step": 4,
                    "note": "public GaussianRasterizer API + torch.autograd.backward (at N > 1 followed by the all-reduce of the six parameter-gradient tensors, 176 MB at config3); per step the camera (view / projection matrices, position) is copied from pinned host memory (double-buffered on a copy stream) and a result scalar is read back.  The Gaussian arrays are resident like model weights and so is the view's supervision (the upstream-gradient planes), as in GScream, whose cameras keep every view's images on the GPU (scene/cameras.py) — the reference arm is timed with everything resident as well.  The rendered planes are NOT copied back: their consumer (the loss) lives on the device.  Round 1 reported the supervision-from-host form under this key; it is kept below"},
            "e2e_supervision_from_host": {"value": e2e_sup_value, "unit": "views/s", "h2d_bytes_per_step": h2d_view_bytes, "d2h_bytes_per_step": 4,
                    "note": "as e2e, but the view's upstream-gradient planes (C colour + depth + uncertainty: 282 MB at config3) are copied from pinned host memory every step as well: a measurement of the PCIe link (~56 GB/s), not of the kernels"},
            "e2e_device_resident_api": {"value": api_value, "unit": "views/s", "ms_per_step": api_ms / args.steps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "public GaussianRasterizer API + torch.autograd.backward with every tensor resident on the device: the like-for-like arm against `--impl reference`, which is timed exactly this way (its `value`)"},
            "e2e_all_inputs_from_host": {"value": e2e_all_value, "unit": "views/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "note": "strictest form: ALL Gaussian arrays, camera and upstream gradient planes copied from pinned host memory every step; PCIe-bound"},
            "gpu_launches": launches, "clocks": clocks,
            "stage_ms": stage_ms, "stage_ms_note": "per-stage CUDA-event means inside the timed region; max over ranks" if world > 1 else "per-stage CUDA-event means inside the timed region",
            "roofline": {"bound": "hbm", "kernel": bwd_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the `ncu --set full` capture recorded
                         # for THIS build of the library (profiles/traffic.json, keyed by the build digest); null for any other build
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": bytes_bwd,
                         "avg_launch_ms": t_bwd, "peak_source": peak_src,
                         "note": "the blend kernels are issue / tensor-pipe bound, not HBM bound: their DRAM traffic is a quarter of the algorithmic bytes (L2 absorbs the re-gathers), see DESIGN.md section 4 and profiles/r2_*"},
        }
        if world > 1:
            out["stage_ms_rank0"] = stage_ms_rank0
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_port(cfg)
    return out


def other_configs():
    """The other BASELINE.json configurations, measured by this same run so that they are driver-run numbers too: config2
    (configs[1]), the config4 substitute (configs[3]) and SURVEY 8d's heavy variant of config3, each with both arms.
    One child process per arm; device timed like the main line; a few seconds each."""
    res = {}
    # (config4: tensor sizes change every iteration — the kept offsets — so torch's caching allocator keeps calling cudaMalloc for
    # the first dozens of iterations, 10-45 ms each; 30 warm-up iterations let its pool settle in both arms)
    for name, wl, steps, warm, extra in (("config2", "config2", 60, 5, []), ("config4", "config4", 60, 30, []),
                                         ("config3_heavy", "config3", 20, 5, ["--scale-mult", "3.0", "--value-only"])):
        for impl in ("ours", "reference"):
            cmd = [sys.executable, os.path.abspath(__file__), "--workload", wl, "--impl", impl, "--steps", str(steps), "--warmup", str(warm),
                   "--no-cpu-baseline", "--no-other-configs"] + extra
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                d = json.loads(r.stdout.strip().splitlines()[-1])
                if "unavailable" in d:
                    res.setdefault(name, {})[impl] = {"unavailable": d["unavailable"]}
                    continue
                e = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"]}
                if "config" in d:
                    e["workload"] = d["config"]["workload"]
                if "num_rendered" in d:
                    e["num_rendered"] = d["num_rendered"]
                if "stage_ms" in d:
                    e["stage_ms"] = d["stage_ms"]
                if "e2e_device_resident_api" in d:
                    e["device_resident_api"] = d["e2e_device_resident_api"]["value"]
                res.setdefault(name, {})[impl] = e
            except Exception as ex:  # a failed extra must not lose the main line
                res.setdefault(name, {})[impl] = {"error": repr(ex)[:200]}
        a, b = res[name].get("ours", {}), res[name].get("reference", {})
        if "value" in a and "value" in b:
            res[name]["ratio"] = a["value"] / b["value"]
    res["config3_heavy"]["note"] = "config3 with the splat scale x3 (SURVEY 8d's 'heavy' variant: ~40 tile instances per Gaussian, as real scenes have); device-resident bucket path vs the reference's device-resident arm"
    return res


def run_reference(args, cfg, rank, world, local):
    """The reference's own CUDA build through its own Python API (rank 0 only)."""
    if rank != 0:
        return None
    import _ref_utils as ru
    from gscream_b200 import scenes
    P, W, H, C, seed = cfg["P"], cfg["W"], cfg["H"], cfg["C"], cfg["seed"]
    if not ru.ref_available(C):
        return {"impl": "reference", "unavailable": "oracle/_ref/dgr%d not built (oracle/build_ref.py needs /root/reference)" % C}
    ref = ru.load_ref(C)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    scene = {k: v.to(dev) for k, v in scenes.make_scene(P, W, H, C, seed, scale_mult=args.scale_mult).items()}
    cam = scenes.make_camera(W, H)
    gc, gd, gu = (g.to(dev) for g in scenes.make_upstream_grads(C, W, H, seed))
    st = ref.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=scene["bg"],
                                           scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev),
                                           sh_degree=1, campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    rast = ref.GaussianRasterizer(raster_settings=st)
    leaves = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "colors", "opacities", "uncertainties", "scales", "rotations")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

    def step():
        color, depth, unc, radii = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], uncertainties=leaves["uncertainties"],
                                        shs=None, colors_precomp=leaves["colors"], scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None)
        torch.autograd.backward((color, depth, unc), (gc, gd, gu))
        for t in list(leaves.values()) + [m2d]:
            t.grad = None

    ms, clocks = _timed(step, args.steps, args.warmup, 1, ClockSampler(local))
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step
    cores = os.cpu_count()
    sample = "full workload, %d steps: the reference has no CPU rasterizer; its CUDA sources (unmodified, NUM_CHANNELS=%d via pre-include) recompiled for sm_100a, 1 GPU, device-resident inputs" % (args.steps, C)
    return {"impl": "reference", "metric": "fwd+bwd views/sec", "value": value, "unit": "views/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded Gaussian cloud, SURVEY.md 8d)",
            "config": {"workload": "%s%s: %d Gaussians, %dx%d, %d feature channels + depth + uncertainty, fwd+bwd" % (
                args.workload, "" if args.scale_mult == 1.0 else " heavy (splat scale x%g)" % args.scale_mult, P, W, H, C)},
            "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clocks": clocks}


def run_train_step(args, rank, world, local):
    """BASELINE.json configs[3] substitute (SURVEY.md 8d "Config 4"): the train-step-equivalent loop of tests/_train_step.py,
    1 GPU.  `--impl ours`: fused CUDA decode + this repo's rasterizer + fused L1/SSIM and aligned-depth losses + fused statistics; `--impl reference`: torch decode (the reference's
    generate_neural_gaussians restated) + the reference's own rasterizer build + eager losses + torch.optim.Adam (its own code path).
    `ours` additionally uses the fused four-scale depth gradient loss and the fused multi-tensor Adam."""
    if rank != 0:
        return None
    import _train_step as ts
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    A, k, W, H = 100000, 10, 1008, 567
    if args.impl == "reference":
        import _ref_utils as ru
        if not ru.ref_available(3):
            return {"impl": "reference", "unavailable": "oracle/_ref/dgr3 not built (oracle/build_ref.py needs /root/reference)"}
        mod, dec = ru.load_ref(3), ts.torch_decode
    else:
        from gscream_b200 import rasterizer as mod
        dec = ts.fused_decode
    loop = ts.TrainStep(mod, dec, A=A, k=k, W=W, H=H, device=dev, fused_losses=(args.impl == "ours"))
    ms, clocks = _timed(loop.step, args.steps, args.warmup, 1, ClockSampler(local))
    ms_per_step = ms / args.steps
    out = {"metric": "train-step-equivalent iters/sec", "value": 1000.0 / ms_per_step, "unit": "iters/s", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic anchors / target image / target depth (the SPIN-NeRF 'book' scene and train.py's dependencies are absent)",
           "config": {"workload": "config4 substitute: %d anchors x %d offsets, %dx%d, prefilter + decode + rasterize (RGB+depth+uncertainty) + "
                                  "L1/SSIM/aligned-depth/4-scale depth-gradient losses + densification statistics + Adam; P = %d Gaussians from %d visible anchors" % (A, k, W, H, loop.last["P"], loop.last["n_vis"]),
                      "decode": "fused CUDA (gsr_decode_*)" if args.impl == "ours" else "torch eager (reference code path)",
                      "rasterizer": "libgsr_b200" if args.impl == "ours" else "reference CUDA build (oracle/_ref/dgr3)",
                      "losses": "fused L1 + SSIM (gsr_l1_ssim_*), aligned-depth L1 (gsr_depth_align_l1_*), 4-scale depth gradient loss (gsr_depth_grad_*)" if args.impl == "ours" else "torch eager",
                      "optimizer": "fused multi-tensor Adam (gsr_adam_step)" if args.impl == "ours" else "torch.optim.Adam (foreach)",
                      "densification_statistics": "fused (gsr_training_statis)" if args.impl == "ours" else "torch eager (reference method)"},
           "clocks": clocks}
    if args.impl == "reference":
        out["impl"] = "reference"
        out["cpu_baseline"] = {"value": out["value"], "unit": "iters/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "full loop, %d steps, reference rasterizer + eager torch decode on the same GPU" % args.steps}
        out["e2e"] = {"value": out["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        return out
    from gscream_b200 import _lib
    lib = _lib.load()
    # ---- where the iteration's GPU time goes: per-stage CUDA-event sums over 10 more iterations (library kernels only) ----
    names = ("preprocess", "depth_order_scan", "binning", "blend_forward", "blend_backward", "gaussian_backward", "decode_forward",
             "decode_backward", "losses_forward", "losses_backward", "adam_step")
    lib.gsr_profile_enable(1)
    for _ in range(10):
        loop.step()
    torch.cuda.synchronize()
    buf = np.zeros(256, np.float32)
    stage_ms = {}
    for sid, name in enumerate(names):
        n = lib.gsr_profile_read(sid, buf.ctypes.data, 256)
        stage_ms[name] = float(buf[:n].sum() / 10.0) if n > 0 else None
    lib.gsr_profile_enable(0)
    out["stage_ms"] = stage_ms
    n_param = sum(p_.numel() for p_ in loop.pc.parameters())
    peak, peak_src = _peaks()
    if stage_ms["adam_step"]:
        adam_bytes = 28 * n_param    # read param, grad, exp_avg, exp_avg_sq; write param, exp_avg, exp_avg_sq (fp32)
        out["optimizer"] = {"kernel": "adam_step_kernel", "parameters": n_param, "kernel_ms": stage_ms["adam_step"],
                            "roofline": {"bound": "hbm", "algorithmic_bytes": adam_bytes, "achieved": adam_bytes / (stage_ms["adam_step"] * 1e-3) / 1e9,
                                         "peak": peak, "unit": "GB/s", "frac": adam_bytes / (stage_ms["adam_step"] * 1e-3) / 1e9 / peak, "peak_source": peak_src}}
    # ---- decode alone: fused CUDA vs eager torch on the same anchors / visibility (fwd + bwd), device time ----
    pc, campos = loop.pc, loop.campos
    with torch.no_grad():
        rast = mod.GaussianRasterizer(raster_settings=loop.settings)
        vis = rast.visible_filter(means3D=pc._anchor, scales=pc.get_scaling[:, :3], rotations=pc.get_rotation, cov3D_precomp=None) > 0

    def decode_step(fn):
        def f():
            outs = fn(campos, pc, vis)
            loss = sum(o.sum() for o in outs[:7])
            for p_ in pc.parameters():
                p_.grad = None
            loss.backward()
        return f
    lib.gsr_profile_enable(1)
    lib.gsr_launch_count(1)
    ms_f, _ = _timed(decode_step(ts.fused_decode), 20, 3, 1)
    launches = int(lib.gsr_launch_count(0))
    buf = np.zeros(256, np.float32)
    st = {}
    for sid, name in ((6, "decode_forward"), (7, "decode_backward")):
        n = lib.gsr_profile_read(sid, buf.ctypes.data, 256)
        # stage 6 is recorded twice per decode (stage 1 and stage 2): report their sum per decode
        st[name] = float(buf[:n].sum() / 23.0) if n > 0 else None
    lib.gsr_profile_enable(0)
    ms_t, _ = _timed(decode_step(ts.torch_decode), 20, 3, 1)
    n_vis, P = loop.last["n_vis"], loop.last["P"]
    fwd_bytes = n_vis * (128 + 12 + 12 * k + 24 + 5 * k) + A + 60 * P
    bwd_bytes = n_vis * (128 + 12 + 12 * k + 24) * 2 + 64 * P + 4 * k * n_vis
    out["decode"] = {"fused_fwd_bwd_ms": ms_f / 20, "torch_fwd_bwd_ms": ms_t / 20, "kernel_ms": st, "gpu_launches_per_decode": launches / 23.0,
                     "visible_anchors": n_vis, "gaussians": P,
                     "roofline": {"bound": "hbm", "kernel": "decode_backward_kernel", "algorithmic_bytes_fwd": fwd_bytes, "algorithmic_bytes_bwd": bwd_bytes,
                                  "achieved_fwd": fwd_bytes / (st["decode_forward"] * 1e-3) / 1e9 if st["decode_forward"] else None,
                                  "achieved_bwd": bwd_bytes / (st["decode_backward"] * 1e-3) / 1e9 if st["decode_backward"] else None,
                                  "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                                  "note": "8.4 kMAC of MLP per anchor against ~0.7 KB: not HBM bound; 3xTF32 mma.sync on the tensor pipe, bounded by warps in flight (profiles/r2_decode_mma.md)"}}
    out["gpu_launches"] = launches
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config2", "config3", "config4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--yaw-deg", type=float, default=None, help="development aid: camera yaw of this rank's view (default: 3 degrees x (rank - (N-1)/2))")
    ap.add_argument("--value-only", action="store_true", help="development aid: skip the end-to-end arms (the printed line then lacks them)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the config2 / config4 arms that the default 1-GPU run appends")
    ap.add_argument("--scale-mult", type=float, default=1.0,
                    help="multiply the synthetic splat scale (SURVEY 8d 'heavy' variant: 3.0); 1.0 is the BASELINE.json workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    from gscream_b200 import scenes
    rank, world, local = _dist_env()
    if args.workload == "config4":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        if rank == 0:
            print(json.dumps(run_train_step(args, rank, world, local)), flush=True)
        return
    cfg = scenes.CONFIGS[args.workload]
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args, cfg, rank, world, local)), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    out = run_ours(args, cfg, rank, world, local)
    if rank == 0 and world == 1 and args.workload == "config3" and args.scale_mult == 1.0 and not args.no_other_configs and not args.value_only:
        torch.cuda.empty_cache()
        out["other_configs"] = other_configs()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
