set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain_mass.py -q -x --tb=short -k cross_check 2>&1 | tail -40 > gpurun_out/exp2_crosscheck.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/exp2_pytest_gpu.log; cat gpurun_out/exp2_pytest_gpu.log
run() { w=$1; shift; o=""; for kv in "$@"; do o="$o --opt $kv"; done
  timeout 600 python bench.py --workload $w --no-cpu $o 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('RUN $w $*', 'ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], d['roofline']['kernels_ms'], d['quality'].get('queue_frac'), d['quality'].get('queue_ipm_iters_mean'))"; }
run cartpole
run cartpole split=1
run cartpole split=3
run cartpole split=4
run evaporation
run evaporation coop=0
run cartpole_bx
run cartpole_bx coop=0
run chain_mass
