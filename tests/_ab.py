"""Scratch A/B timing of library variants: python tests/_ab.py <config> <lib1> <lib2> ..."""
import os, subprocess, sys, json
cfg = sys.argv[1]
for lib in sys.argv[2:]:
    env = dict(os.environ)
    if lib != "default": env["GSR_LIB_PATH"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, "bench.py", "--workload", cfg, "--steps", "40", "--warmup", "5", "--no-cpu-baseline", "--no-other-configs"] + os.environ.get("AB_EXTRA", "").split(), env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(lib, "value %.1f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["stage_ms"].items()}, flush=True)
    except Exception as e:
        print(lib, "FAILED", out.stdout[-500:], out.stderr[-1500:])
