"""bench.py's reference arm runs on CPU: its JSON line must carry the contract's keys (the GPU arm is exercised on the
GPU box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-n", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gvoxels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "1348128 faces" in d["config"]["workload"] and "1024^3" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--ref-n", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_benchmarks_runner_parses_the_cli_labels():
    """tools/benchmarks.py: sub-scope lines are summed into the record their main line closes; names in snake case."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import importlib
    from collections import defaultdict
    bm = importlib.import_module("benchmarks")
    assert bm.snake("B200Vox::Memory") == "b200_vox__memory"
    assert bm.snake("SequentialJFA") == "sequential_jfa"
    table = defaultdict(lambda: defaultdict(list))
    text = "\n".join(["[B200Vox::Memory]: 1.5 ms", "[B200Vox::Processing]: 0.25 ms", "[B200Vox(bunny)]: 2 ms",
                      "noise", "[B200Vox::Memory]: 1 ms", "[B200Vox::Processing]: 0.5 ms", "[B200Vox(bunny)]: 1.75 ms",
                      "[B200JFA::Memory]: 3 ms", "[B200JFA::Processing]: 4 ms", "[B200JFA]: 7.5 ms"])
    bm.parse_run(text, "64", table)
    assert table["b200_vox"]["64"] == [{"b200_vox__memory": 1.5, "b200_vox__processing": 0.25, "b200_vox": 2.0},
                                       {"b200_vox__memory": 1.0, "b200_vox__processing": 0.5, "b200_vox": 1.75}]
    assert table["b200_jfa"]["64"] == [{"b200_jfa__memory": 3.0, "b200_jfa__processing": 4.0, "b200_jfa": 7.5}]
