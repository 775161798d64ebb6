#!/bin/bash
# tests + bench + launch list after a fix-up (tier 2/3) change
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() {
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $1 2>> gpurun_out/err_fix.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fixup_kernel_ms'), d.get('gpu_launches'))"
}
run ""
run "--holes 0"
run "--workload chain2048"
timeout 300 python tools/dev_scale_check.py 2>&1 | tail -6
for h in 0.01 0; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 8 --csv --log-file gpurun_out/launches_h$h.csv \
    python bench.py --steps 3 --warmup 1 --holes $h --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_h$h.log 2>&1
done
python - <<'PY'
import csv
for h in ("0.01","0"):
    rows=[r for r in csv.reader(open(f"gpurun_out/launches_h{h}.csv")) if len(r)>5 and r[0].isdigit()]
    print("holes",h)
    for r in rows[-4:]: print("  ", r[4][:60], r[-1])
PY
tail -3 gpurun_out/err_fix.log
