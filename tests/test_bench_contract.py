"""bench.py's reference arm runs on CPU: check the JSON line the driver parses (keys, units, the cpu_baseline and e2e
objects of the reference arm) without a GPU; and that the device arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "scores/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["config"]["workload"].startswith("C3 dyadic all-pairs")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_ranks_other_than_zero_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_device_arm_does_not_fall_back_to_the_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "scores/s" not in r.stdout


def test_smoke_tie_check_accepts_fp32_ties_and_rejects_real_differences():
    """The index-set check of __graft_entry__.smoke(): a swap inside an fp32 tie at the k-th distance passes
    (and leaves smoke()'s helpers callable), a swap with a clearly worse candidate fails."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    D = np.array([[0.1, 0.2, 0.3, 0.3 * (1 + 1e-7), 0.9]])
    want_idx, want_val = np.array([[0, 1, 2]]), np.array([[0.1, 0.2, 0.3]])
    ge.check_topk_sets(D, want_idx, want_val, np.array([[0, 1, 3]]))          # tie: candidate 3 for 2
    ge.check_topk_sets(D, want_idx, want_val, np.array([[0, 1, 2]]))
    with pytest.raises(AssertionError):
        ge.check_topk_sets(D, want_idx, want_val, np.array([[0, 1, 4]]))
