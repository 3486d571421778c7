#!/bin/bash
# experiment driver (run under gpurun): bench the score kernel under a few knobs
set -x
B='python bench.py --steps 10 --warmup 3'
show() { python -c "
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(b['value']/1e9,1), 'G/s step', round(b['ms_per_step'],3), 'kernel', b['roofline']['kernel_ms'])" $1 "$2"; }
$B > /tmp/a.json 2>/dev/null; show /tmp/a.json "nepi12 stride32"
CFL_SCORE_SAMPLE_STRIDE=16 $B > /tmp/b.json 2>/dev/null; show /tmp/b.json "nepi12 stride16"
CFL_SCORE_SAMPLE_STRIDE=8 $B > /tmp/c.json 2>/dev/null; show /tmp/c.json "nepi12 stride8"
CFL_NVCC_EXTRA="-DCFL_SU_NEPI=8" python compatibility-family-learning_b200/build.py --force > /dev/null 2>&1
$B > /tmp/d.json 2>/dev/null; show /tmp/d.json "nepi8 stride32"
CFL_SCORE_SAMPLE_STRIDE=16 $B > /tmp/e.json 2>/dev/null; show /tmp/e.json "nepi8 stride16"
CFL_NVCC_EXTRA="-DCFL_SU_NEPI=16" python compatibility-family-learning_b200/build.py --force > /dev/null 2>&1
grep -A1 "score_umma_kernelILi3" compatibility-family-learning_b200/build/score_umma.log | grep -E "spill|Used" | head -3
CFL_SCORE_SAMPLE_STRIDE=16 $B > /tmp/f.json 2>/dev/null; show /tmp/f.json "nepi16 stride16"
