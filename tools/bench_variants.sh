# Tuning helper: time bench.py against every mpc4rl_b200/variants_*.so (launch-bound / block-size variants).
for so in mpc4rl_b200/variants_*.so; do
  echo "== $so"
  RLMPC_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value']))
    elif l: print(l[:200])
"
done
