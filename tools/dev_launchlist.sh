#!/bin/bash
# ncu per-launch durations (cold, serialised) of our kernels, with and without holes, and a full capture of the tier-3 kernel
mkdir -p gpurun_out
for h in 0.01 0; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 12 --csv --log-file gpurun_out/launches_h$h.csv \
    python bench.py --steps 3 --warmup 1 --holes $h --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_h$h.log 2>&1
done
python - <<'PY'
import csv
for h in ("0.01","0"):
    rows=[r for r in csv.reader(open(f"gpurun_out/launches_h{h}.csv")) if len(r)>5 and r[0].isdigit()]
    print("holes",h)
    for r in rows[-6:]: print("  ", r[4][:60], r[-1])
PY
ncu --set full --clock-control none --import-source on -k regex:k_fixup_cells -s 2 -c 1 -f -o gpurun_out/prof_t3 \
    python bench.py --steps 2 --warmup 3 --holes 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_t3.log 2>&1
tail -2 gpurun_out/ncu_t3.log
