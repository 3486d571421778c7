# pass C (score_lb_kernel) under the CFL_SCORE_DBG_MODE experiments: what bounds the kernel
#   0 normal | 1 epilogue does nothing | 3 MMA only (no TMA) | 4 epilogue = tcgen05.ld + wait | 8 = ld + bound, no pushes
for m in 0 1 3 4 8; do
  CFL_SCORE_DBG_MODE=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('mode $m kernel_ms', b['roofline']['kernel_ms'], 'step', b['ms_per_step'])"
done
