# the cart-pole multi-GPU lines only (final build): gpurun --gpus N -- bash tools/r02z_multi_gpu_short.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --no-cpu 2>/dev/null | tail -1 > gpurun_out/r02z_bench_cartpole_weak_${N}gpu.json
$TR bench.py --gpus $N --no-cpu --scaling strong 2>/dev/null | tail -1 > gpurun_out/r02z_bench_cartpole_strong_${N}gpu.json
$TR bench.py --gpus $N --no-cpu --scaling strong --graph 2>/dev/null | tail -1 > gpurun_out/r02z_bench_cartpole_strong_graph_${N}gpu.json
for f in gpurun_out/r02z_bench_cartpole_*_${N}gpu.json; do python - $f <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "allreduce_ms %.4f" % d["allreduce_ms"], d["roofline"]["kernels_ms"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
