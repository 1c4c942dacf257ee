# traversal occupancy sweep: resident blocks per SM (register budget 126 / 96 / 80 per thread)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 5 6 7 8; do
  TMD_NL_MINB=$m python bench.py --quick --steps 60 --warmup 10 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('MINB=$m', 'value=%.4g' % d['value'], 'ms_per_step=%.4f' % d['ms_per_step'], 'force_eval_ms=%.4f' % d['roofline']['launch_ms'], 'frac=%.3f' % d['roofline']['frac'], d['config']['nlist'])"
done
