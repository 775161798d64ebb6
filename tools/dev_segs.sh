#!/bin/bash
# calibration of the fused kernel's work-queue levels (TE_FUSED_SEGS): one bench line per setting
mkdir -p gpurun_out
run() {
  TE_FUSED_SEGS="$1" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e ${2:-} 2>> gpurun_out/err_segs.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fixup_kernel_ms'))"
}
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run ""
run "" "--holes 0"
run "160:0.8,48:0.15,16:0.05"
run "" "--rows 4096 --cols 4096"
run "" "--workload chain2048"
run "" "--workload batched512"
timeout 300 python tools/dev_scale_check.py 2>&1 | tail -6
tail -3 gpurun_out/err_segs.log
