for k in 0.2 0.25 0.3 0.35 0.4; do
  python bench.py --quick --steps 100 --warmup 10 --skin $k 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('skin=$k', 'ms_per_step=%.4f' % d['ms_per_step'], 'force_eval_ms=%.4f' % d['roofline']['launch_ms'], d['config']['nlist'])"
done
