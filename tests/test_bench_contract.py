"""CPU: the reference arm of bench.py prints ONE JSON line with the driver's contract keys (the GPU arm needs a B200)."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-n", "16", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "vcycle_gdof_per_s" and d["unit"] == "GDOF/s"
    assert d["steps"] == 2 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert "workload" in d["config"] and "256^3" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--ref-n", "16"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_both_arms_print_the_same_config_and_geometry_helpers():
    """the reference arm runs the SAME workload: its `config` object is built by the function the GPU arm uses; the
    slab geometry helpers keep n^3 DOF per rank and give the 512^3 cube of BASELINE config 5 at n = 256 on 8 ranks"""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c1 = bench.common_config(256, 1)
    assert c1["dof"] == 256 ** 3 and "256^3" in c1["workload"] and "zslab" not in c1["workload"]
    c8 = bench.common_config(256, 8, "cube")
    assert c8["dof"] == 8 * 256 ** 3 and c8["workload"].endswith("global_512x512x512")
    assert bench.common_config(256, 8, "slab")["workload"].endswith("global_256x256x2048")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-n", "12", "--steps", "1",
                          "--warmup", "1", "--gpus", "8"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="8", LOCAL_RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["config"] == c8 and d["n_gpus"] == 8
    from mlamg import distributed as md
    for n, world in ((256, 8), (48, 4), (16, 2)):
        nx, ny, nzl = md.slab_geometry(n, world, "cube")
        assert nx * ny * nzl == n ** 3 and (nx, ny) == (2 * n, 2 * n)
        assert md.slab_geometry(n, world, "slab") == (n, n, n)
        lam = md.slab_lambda_max(n, world, "cube")
        assert abs(lam - (1 + (2 * np.cos(np.pi / (2 * n + 1)) + np.cos(np.pi / (nzl * world + 1))) / 3)) < 1e-15
    assert md.slab_geometry(256, 8, "cube") == (512, 512, 64) and md.slab_geometry(256, 1, "cube") == (256, 256, 256)
