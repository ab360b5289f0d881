# last GPU pass of round 2: parity suite, smoke, launch-shape variants on the headline workload, default bench line
# (gpurun --timeout 300 -- bash tools/r02_final_check.sh)
mkdir -p gpurun_out
T=r02f
timeout 120 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 > gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_smoke.log
run() {
  timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['kernels_ms'])"
}
{
echo "== default"; run
echo "== split=1"; run --opt split=1
echo "== split=3"; run --opt split=3
echo "== split=4"; run --opt split=4
echo "== RLMPC_COOP_BLOCKS_PER_SM=3"; RLMPC_COOP_BLOCKS_PER_SM=3 run
echo "== RLMPC_COOP_BLOCKS_PER_SM=3 split=3"; RLMPC_COOP_BLOCKS_PER_SM=3 run --opt split=3
} > gpurun_out/${T}_variants_split_headline.log 2>&1
cat gpurun_out/${T}_variants_split_headline.log
