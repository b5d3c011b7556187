"""The reference arm of bench.py (`--impl reference`) runs on host cores only -- check the JSON contract here on the CPU box;
and the product arm must refuse to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--no-posenet")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dcnv3_fwd_bwd_algorithmic_GBps" and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    from baseline import reference as R
    # the reference's own dcnv3_func.py when it is staged (baseline/_ref, written by build()) or present; the oracle port otherwise
    assert line["cpu_baseline"]["kind"] == ("reference" if R.root() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "N=64" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 1 and "workload" in line["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config(1, "T")        # the SAME config our arm prints (same_config in the driver's ratio)


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
