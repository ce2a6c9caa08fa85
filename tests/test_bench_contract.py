"""bench.py contract on CPU: the reference arm (the unmodified reference FCGANModel from baseline/_ref or /root/reference on
the host cores, else the oracle port) must print ONE JSON line with the keys the driver reads, on the same metric / unit /
config as the GPU arm, using every host core even when torchrun exported OMP_NUM_THREADS=1."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what torchrun exports: the arm must override it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--batch", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    import bench
    import argparse
    assert d["impl"] == "reference" and d["metric"] == bench.WORKLOADS["fcgan"]["metric"] and d["unit"] == bench.UNIT
    assert d["config"] == bench.line_config(argparse.Namespace(config="fcgan", pool_size=50, batch=1), 1)   # what our arm prints
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
