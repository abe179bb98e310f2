#!/usr/bin/env python
"""Build experimental variants of libvpb200.so (compile-time tunables) and time the JFA passes of each on the GPU.
   python tools/variants.py build            (here, no GPU)
   python tools/variants.py run [n]          (on the GPU box)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "cuda_mesh_voxelization_b200", "build", "variants")
VARIANTS = {
    "base": [],
    "mb2": ["-DVPB_FLOOD_MINBLOCKS=2"],
    "mb4": ["-DVPB_FLOOD_MINBLOCKS=4"],
    "lz64": ["-DVPB_FLOOD_LZ=64"],
    "mb4_lz64": ["-DVPB_FLOOD_MINBLOCKS=4", "-DVPB_FLOOD_LZ=64"],
}
if sys.argv[1] == "build":
    from cuda_mesh_voxelization_b200 import _build
    os.makedirs(VDIR, exist_ok=True)
    for name, flags in VARIANTS.items():
        print(name, _build.build(force=True, extra_flags=flags, out=os.path.join(VDIR, f"libvpb200_{name}.so")))
elif sys.argv[1] == "run":
    n = sys.argv[2] if len(sys.argv) > 2 else "512"
    for name in VARIANTS:
        env = dict(os.environ, VPB_LIB=os.path.join(VDIR, f"libvpb200_{name}.so"))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", n, "--steps", "3", "--warmup", "2", "--no-cpu-baseline"],
                             env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(name, "ms/step %.2f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["ms_per_pass_by_k"].items()}, flush=True)
        except Exception as e:
            print(name, "FAILED", e, out.stderr[-400:])
