#!/bin/bash
for cfg in "8 " "9 224" "10 200" "10 192" "12 168"; do set -- $cfg
  make -C traversability_estimation_b200/csrc -B TE_WPC=$1 TE_REGS=$2 -s 2>&1 | grep -E "error"
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('wpc=$1 regs=$2', round(d['value']), d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'])"
done
