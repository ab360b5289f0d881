# A/B of library builds on one workload: gpurun -- bash tools/r02b_variants.sh <workload> [bench flags]
W=${1:-cartpole}; shift
run() { python bench.py --workload $W --warmup 3 --no-cpu --steps 10 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], d['roofline']['kernels_ms'], d['quality'].get('queue_frac'), d['quality'].get('queue_ipm_iters_mean'), d['quality'].get('status0_frac_last_step'))"; }
echo "== default build"; run "$@"
for so in mpc4rl_b200/variants_*.so; do
  [ -e "$so" ] || continue
  echo "== $so"; RLMPC_B200_LIB=$PWD/$so run "$@"
done
