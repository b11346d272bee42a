"""CPU-side checks of bench.py: the algorithmic-byte model equals SURVEY.md 8(d)'s figures, the
host-to-host pipeline's sample groups partition the batch, the reference arm prints the contract's
JSON line, and our arm refuses to run without a CUDA device (no CPU fallback)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest
import torch

from e2e_parking_carla_b200.synthetic import LiftSplatShape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_equal_survey_figures(bench):
    """cfg1/2 fp32 fwd 12.07 MB/sample = 1.049 + 0.786 + 10.240, bwd 13.91 MB;
    bf16 inputs 11.16 / 12.08; cfg4 fp32 44.89 / 48.82 MB/sample."""
    ab = bench.algorithmic_bytes(LiftSplatShape(batch=16, channels=64), 4)
    assert ab["fwd"] == 4 * 1024 * (64 + 48) * 4 + 64 * 200 * 200 * 4 == 12_075_008
    assert ab["bwd"] == 64 * 200 * 200 * 4 + 2 * 4 * 1024 * (64 + 48) * 4 == 13_910_016
    assert (ab["fwd"] + ab["bwd"]) * 16 == 415_760_384          # the step the bench line is quoted on
    # the per-kernel figures add up to the step's
    assert ab["splat_fwd"] == ab["fwd"] and ab["bwd_transpose"] + ab["bwd_gather"] == ab["bwd"]
    half = bench.algorithmic_bytes(LiftSplatShape(batch=16, channels=64), 2)
    assert round(half["fwd"] / 1e6, 2) == 11.16 and round(half["bwd"] / 1e6, 2) == 12.08
    stress = bench.algorithmic_bytes(LiftSplatShape.stress(batch=32), 4)
    assert round(stress["fwd"] / 1e6, 2) == 44.89 and round(stress["bwd"] / 1e6, 2) == 48.82


@pytest.mark.parametrize("batch", [1, 2, 3, 4, 12, 16, 32, 33])
@pytest.mark.parametrize("chunks", [0, 1, 3, 4])
def test_e2e_groups_partition_the_batch(bench, batch, chunks):
    sizes = bench._e2e_sizes(batch, chunks)
    assert sum(sizes) == batch and all(s > 0 for s in sizes)
    if chunks == 0 and batch >= 4:
        assert len(sizes) == 3 and sizes[0] == sizes[2] <= sizes[1]     # small group at each end


def test_our_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr and res.stdout.strip() == ""


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: one JSON line, same metric / unit / config as our arm,
    impl = reference, a cpu_baseline describing the run and an e2e that repeats the value with no
    copies.  One sample of the workload keeps it to a few seconds here."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "lift_splat_fwd_bwd_samples_per_s" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "configs[1]" in line["config"]["workload"] and line["config"]["reference_batch"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    assert cb["kind"] == ("reference" if os.path.isdir("/root/reference") else "port")
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
    assert abs(line["value"] - 1000.0 / line["ms_per_step"]) < 1e-6 * line["value"]


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
