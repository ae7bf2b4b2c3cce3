"""bench.py host-side checks that need no GPU: the reference arm prints one well-formed JSON line,
the synthetic data is keyed by the global utterance index, and no time-bounded loop of the GPU arm
contains a collective (that dead-locks multi-rank runs)."""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cells/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_non_zero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_synthetic_inputs_are_keyed_by_global_index():
    sys.path.insert(0, ROOT)
    from bench import synthetic_numpy
    a, _ = synthetic_numpy(0, 4, 6, 5)
    b, _ = synthetic_numpy(2, 2, 6, 5)
    assert np.array_equal(a[2:], b)


def test_no_time_bounded_loop_in_the_gpu_arm():
    """Every forward_backward call advances the loss exchange's call counter, which must stay in step across
    ranks: the GPU arm may not contain a loop whose trip count depends on the wall clock."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    gpu_arm = src[src.index("def run_b200"):src.index("def main")]
    assert "while time.perf_counter()" not in gpu_arm


def test_graph_length_tiles_any_step_count():
    sys.path.insert(0, ROOT)
    from bench import graph_len
    for steps in (1, 2, 7, 20, 23, 200):
        g = graph_len(steps)
        assert 1 <= g <= 10 and steps % g == 0
    assert graph_len(20) == 10 and graph_len(200) == 10 and graph_len(23) == 1


def test_reference_arm_times_the_global_batch():
    env = dict(os.environ, RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--gpus", "2", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["config"]["global_batch"] == 2 and line["config"]["n_batches"] == 2
