for m in 0 1 2 3; do
  CFL_SCORE_DBG_MODE=$m timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('mode $m kernel_ms', b['roofline']['kernel_ms'], 'step', b['ms_per_step'])"
done
