#!/usr/bin/env python
"""Build experimental variants of libvpb200.so (compile-time tunables) and time the JFA passes of each on the GPU.
   python tools/variants.py build            (here, no GPU)
   python tools/variants.py run [n]          (on the GPU box)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "cuda_mesh_voxelization_b200", "build", "variants")
VARIANTS = {
    "f5_r2": ["-DVPB_F5_RPT=2"],
    "f5_r4": ["-DVPB_F5_RPT=4"],
    "f5_r4_c3": ["-DVPB_F5_RPT=4", "-DVPB_F5_CTAS=3"],
}
if os.environ.get("VPB_VARIANT_FLAGS"):
    VARIANTS = {k: v + os.environ["VPB_VARIANT_FLAGS"].split() for k, v in VARIANTS.items()}
if sys.argv[1] == "build":
    from cuda_mesh_voxelization_b200 import _build
    os.makedirs(VDIR, exist_ok=True)
    for name, flags in VARIANTS.items():
        print(name, _build.build(force=True, extra_flags=flags, out=os.path.join(VDIR, f"libvpb200_{name}.so")))
elif sys.argv[1] == "run":
    n = sys.argv[2] if len(sys.argv) > 2 else "512"
    tests = os.environ.get("VPB_VARIANT_TESTS")
    for name in VARIANTS:
        if tests:
            env = dict(os.environ, VPB_LIB=os.path.join(VDIR, f"libvpb200_{name}.so"))
            r = subprocess.run(["timeout", "900", sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q", "-k", tests],
                               env=env, capture_output=True, text=True, cwd=ROOT)
            print(name, "pytest:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
            if r.returncode != 0:
                print(r.stdout[-3000:], flush=True)
        env = dict(os.environ, VPB_LIB=os.path.join(VDIR, f"libvpb200_{name}.so"))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", n, "--steps", "3", "--warmup", "3", "--no-cpu-baseline"],
                             env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(name, "ms/step %.2f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["ms_per_pass_by_k"].items()},
                  "parity", d.get("parity", {}).get("status"), flush=True)
        except Exception as e:
            print(name, "FAILED", e, out.stderr[-400:])
