"""The numerical argument behind the fp16 operand planes of the lower-bound pass, checked on the CPU by the float64
emulation of tools/lb_precision_study.py: (1) no variant's bound ever exceeds the true distance (the rigorous error
terms hold), (2) fp16 operands with fp32 accumulation keep exactly the survivors of the tf32 planes (same 11-bit
significand, values inside the fp16 range), (3) bf16 operands would not."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lower_bound_precision_study_supports_the_fp16_default():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lb_precision_study.py"), "30000", "6"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = {j["variant"]: j for j in (json.loads(l) for l in r.stdout.splitlines() if l.startswith("{"))}
    assert len(rows) == 5
    for j in rows.values():
        assert j["bound_violations"] == 0, j
        assert j["worst_gram_error_over_pe"] <= j["rigorous_u"], j          # the rigorous u really bounds the error
    tf32 = rows["tf32 operands, fp32 accumulate (fallback planes)"]
    fp16 = rows["fp16 operands, fp32 accumulate (shipped)"]
    bf16 = rows["bf16 operands, fp32 accumulate"]
    assert fp16["survivors_per_query"] == tf32["survivors_per_query"]
    assert bf16["survivors_per_query"] > 1.1 * tf32["survivors_per_query"]
