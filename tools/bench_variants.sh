# Tuning helper: time bench.py with option sweeps and against every mpc4rl_b200/variants_*.so.
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['kernels_ms'], d['quality']['status0_frac_last_step'], d['quality'].get('queue_frac'), d['quality'].get('queue_ipm_iters_mean'))
    elif l: print(l[:300])
"
}
echo "== default"; run
for o in "$@"; do echo "== opt $o"; run --opt $o; done
for so in mpc4rl_b200/variants_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"
  RLMPC_B200_LIB=$PWD/$so run
done
