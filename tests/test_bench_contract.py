"""CPU: the reference arm of bench.py (the unmodified reference on the host cores) prints ONE JSON line with the
keys the driver reads; the CUDA arm fails loudly without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_one_contract_line(native_built):
    ref = os.path.join(ROOT, "oracle", "_ref", "kripke_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/kripke_ref is only built where /root/reference exists")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "config1", "--ref-zones", "4,4,4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["metric"] == "grind_time" and d["higher_is_better"] is False
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # honesty: the line says which command line the reference really ran and that it was a sample, not the named block
    assert "--zones 4,4,4" in d["command"] and d["sample_zones"] == [4, 4, 4] and d["sample_is_full_workload"] is False
    assert "--zones" not in d["config"]["workload"]


def test_reference_sample_is_sized_by_host_memory_and_time():
    sys.path.insert(0, ROOT)
    import bench
    z = bench.reference_sample_zones("config2", 20, 5)
    zones, groups, dirs = bench.WORKLOADS["config2"][:3]
    assert all(a <= b for a, b in zip(z, zones)) and z[0] * z[1] * z[2] >= 16 ** 3
    unknowns = z[0] * z[1] * z[2] * groups * dirs
    assert unknowns * 18.5 <= 0.9 * bench.host_mem_available_gb() * 1e9 or z == (16, 16, 16)
    assert bench.reference_sample_zones("config1", 2, 1) == (16, 16, 16)


def test_reference_arm_other_ranks_exit_quietly(native_built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu(native_built):
    import kripke_b200 as kb
    if kb.have_gpu():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--workload", "config1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert "grind_time" not in out.stdout
