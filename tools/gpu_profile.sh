# ncu evidence for profiles/: per-launch time + DRAM bytes of the steady-state RTI steps, and one
# --set full capture of every kernel of one step.  Run under gpurun (1 GPU).
set -x
mkdir -p gpurun_out
# (split=1: the kernels of ONE chain over the whole batch; the default runs two half-batch chains on two streams)
# launch list of the whole run (setup solve, warm-up, timed steps, per-kernel timing pass, e2e pass)
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 4 --warmup 3 --no-cpu --opt split=1 > gpurun_out/b_ncu_l.log 2>&1
# keep the last 4 RTI calls of the device-resident pass (k_begin .. k_sens_sweep), and find how many
# launches of the step kernels precede the first of them
python - <<'PY' > gpurun_out/skip.txt
import csv, re
rows = [r for r in csv.reader(open("gpurun_out/launches_all.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
body = rows[hdr + 1:]
ki = rows[hdr].index("Kernel Name"); mi = rows[hdr].index("Metric Name")
ids = []; names = {}
for r in body:
    if r[0] not in names:
        names[r[0]] = re.sub(r"^void <unnamed>::|<.*", "", r[ki]); ids.append(r[0])
seq = [names[i] for i in ids]
# RTI call = k_begin followed by exactly one k_lin before the next k_begin
starts = [j for j, n in enumerate(seq) if n == "k_begin"]
rti = [s for s, e in zip(starts, starts[1:] + [len(seq)]) if seq[s:e].count("k_lin") == 1 and "k_sens_sweep" in seq[s:e]]
first = rti[2]  # third RTI call: past the warm-up of the caches
last = rti[6] if len(rti) > 6 else len(seq)
step = re.compile(r"k_lin|k_qp1|k_qp3|k_gather|k_condense|k_qp2|k_sens_stage|k_sens_sweep")
print(sum(1 for n in seq[:first] if step.match(n)), sum(1 for n in seq[first:rti[3]] if step.match(n)))
keep = set(ids[first:last])
with open("gpurun_out/launches_rti.csv", "w") as f:
    w = csv.writer(f); w.writerow(rows[hdr])
    for r in body:
        if r[0] in keep: w.writerow(r)
PY
read SKIP COUNT < gpurun_out/skip.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_lin|k_qp1|k_qp3|k_gather|k_condense|k_qp2|k_sens_stage|k_sens_sweep" -s $SKIP -c $COUNT \
  -o gpurun_out/prof_step python bench.py --steps 4 --warmup 3 --no-cpu --opt split=1 > gpurun_out/b_ncu_f.log 2>&1
tail -3 gpurun_out/b_ncu_f.log
