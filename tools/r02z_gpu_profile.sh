# Final ncu / sanitizer / bench pass of round 2 (gpurun -- bash tools/r02z_gpu_profile.sh); outputs in gpurun_out/r02z_*,
# the summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
T=r02z
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_nvidia_smi.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_smoke.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_golden_large.py tests/test_gpu_rti_oracle.py -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_memcheck_new_paths.log; tail -3 gpurun_out/${T}_memcheck_new_paths.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_rti_oracle.py -m gpu -q -k queue_path 2>&1 | tail -6 > gpurun_out/${T}_racecheck_k_qp3.log; tail -3 gpurun_out/${T}_racecheck_k_qp3.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor_op_dmma.sum
for w in cartpole cartpole_tiny_pert evaporation chain_mass; do
  timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_step_$w.csv python tools/profile_step.py --workload $w --steps 2 > gpurun_out/ncu_step_$w.log 2>&1
  tail -1 gpurun_out/ncu_step_$w.log
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_qp3|k_sens_stage|k_sens_sweep|k_qp1|k_lin" -c 10 -o gpurun_out/${T}_full_cartpole python tools/profile_step.py --workload cartpole --steps 1 > gpurun_out/ncu_full_cartpole.log 2>&1
for w in cartpole cartpole_tiny_pert cartpole_replay cartpole_bx evaporation chain_mass chain_mass_6; do
  extra="--no-cpu"; if [ $w = cartpole ] || [ $w = chain_mass ]; then extra=""; fi
  timeout 600 python bench.py --workload $w $extra 2> gpurun_out/${T}_bench_$w.err | tail -1 > gpurun_out/${T}_bench_$w.json
  python - gpurun_out/${T}_bench_$w.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["roofline"]["kernels_ms"], "conv %.1f ms" % d["converge"]["ms"], d["quality"], d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline_literal", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
ls -la gpurun_out | tail -30
