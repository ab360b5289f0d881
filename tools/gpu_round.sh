# One GPU round: tests, smoke, bench, ncu launch list (+ full capture with "full").  Run under gpurun.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -s 40 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
if [ "$1" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_qp1|k_sens|k_lin|k_qp2" -s 30 -c 5 -o gpurun_out/prof_unit python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
fi
for f in gpurun_out/pytest.log gpurun_out/smoke.log gpurun_out/bench.log; do tail -n 5 $f; done
