run() { env "$@" timeout 200 python bench.py --steps 20 --warmup 4 2>/dev/null | tail -1 | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('$*', round(b['value']/1e9,1), 'G/s step', round(b['ms_per_step'],3), 'kernel', b['roofline']['kernel_ms'])"; env "$@" CFL_SCORE_DEBUG=1 timeout 200 python bench.py --steps 2 --warmup 1 2>&1 | grep -a "cfl score" | tail -1; }
run CFL_SCORE_SAMPLE_STRIDE=32 CFL_SCORE_OPT_MULT=4
run CFL_SCORE_SAMPLE_STRIDE=32 CFL_SCORE_OPT_MULT=6
run CFL_SCORE_SAMPLE_STRIDE=48 CFL_SCORE_OPT_MULT=6
run CFL_SCORE_SAMPLE_STRIDE=64 CFL_SCORE_OPT_MULT=6
run CFL_SCORE_SAMPLE_STRIDE=64 CFL_SCORE_OPT_MULT=8
run CFL_SCORE_SAMPLE_STRIDE=16 CFL_SCORE_OPT_MULT=3
