# ncu passes of round 2 (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are copied to profiles/)
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum
for w in cartpole cartpole_tiny_pert chain_mass; do
  timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_step_$w.csv python tools/profile_step.py --workload $w --steps 2 > gpurun_out/ncu_step_$w.log 2>&1
  tail -1 gpurun_out/ncu_step_$w.log
done
