"""CPU test of bench.py's reference arm (`--impl reference`): the driver runs it on the GPU box before the GPU arm and
computes the headline ratio from its line, so the line's keys and units are checked here, where it can run (it needs no GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_line_with_the_contract_keys():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "HMC draws/sec (chains x iters, d=128)" and d["unit"] == "draws/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C2: mcmc::hmc, iso-Gaussian d=128, 4096 chains/GPU, L=10")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "chains x 1100 draws" in cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 1e3 < d["value"] < 1e8 and d["ms_per_step"] > 0   # a CPU rate, not a GPU one


def test_reference_arm_other_ranks_exit_without_work():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device the product arm exits non-zero and prints no result line."""
    import ctypes

    try:
        have_gpu = ctypes.CDLL("libcuda.so.1").cuInit(0) == 0
    except OSError:
        have_gpu = False
    if have_gpu:
        import pytest

        pytest.skip("a CUDA device is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
