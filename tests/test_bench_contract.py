"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with the
required keys, the roofline byte model matches SURVEY.md section 8(d), the pass decomposition is right."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--iters", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "cell_steps_per_sec" and d["unit"] == "cell-steps/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and d["vs_baseline"] is None
    assert d["value"] > 1e5


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_byte_model_and_pass_decomposition():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.step_bytes(15, 20) == 1168 and bench.step_bytes(80, 80) == 4948 and bench.step_bytes(20, 20, smooth=False) == 1340
    assert list(bench._passes(80, 8)) == [8] * 10
    assert list(bench._passes(15, 8)) == [8, 4, 2, 1]
    assert list(bench._passes(20, 8)) == [8, 8, 4]
    assert sum(bench._passes(43, 4)) == 43
