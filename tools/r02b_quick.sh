# quick GPU pass: parity suite + the bench lines of the main workloads (gpurun -- bash tools/r02b_quick.sh [tag])
set -x
TAG=${1:-r02b}
mkdir -p gpurun_out
[ -n "$SKIPTESTS" ] || timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
for w in ${WORKLOADS:-cartpole cartpole_tiny_pert cartpole_replay cartpole_bx evaporation chain_mass}; do
  timeout 600 python bench.py --workload $w --no-cpu 2> gpurun_out/${TAG}_bench_$w.err | tail -1 > gpurun_out/${TAG}_bench_$w.json
  python - gpurun_out/${TAG}_bench_$w.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["roofline"]["kernels_ms"], "conv %.1f ms" % d["converge"]["ms"], d["quality"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
# queue-kernel variants on the headline workload (thread-per-sample queue path instead of the warp-per-sample one)
for v in "coop=0" "coop=0 condense=0" "coop=0 ring=0 condense=0"; do
  o=""; for kv in $v; do o="$o --opt $kv"; done
  timeout 600 python bench.py --workload cartpole --no-cpu $o 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('VARIANT $v', 'ms/step %.3f' % d['ms_per_step'], d['roofline']['kernels_ms'], d['quality'].get('queue_frac'), d['quality'].get('queue_ipm_iters_mean'))"
done
