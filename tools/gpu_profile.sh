# ncu evidence for profiles/: per-launch time + DRAM bytes of the steady-state RTI steps, and one
# --set full capture of every kernel of one step.  Run under gpurun (1 GPU).
set -x
mkdir -p gpurun_out
# launch list: skip the setup solve (61 SQP rounds x 6 kernels + counters) and the warm-up steps, take ~4 steps
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"k_" -s 400 -c 44 --csv --log-file gpurun_out/launches_rti.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/b_ncu_l.log 2>&1
# full capture: kernels of the first timed step (61 x 5 matching launches in the setup, 3 x 7 in the warm-up)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_lin|k_qp1|k_gather|k_condense|k_qp2|k_sens_stage|k_sens_sweep" -s 326 -c 7 \
  -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu_f.log 2>&1
tail -3 gpurun_out/b_ncu_f.log
