#!/bin/bash
# Round-2 A/B session: time the variant builds of the fused kernel, check the candidates against the literal kernel on the whole
# bench map, run the GPU test suite against the product library and against the candidate.
mkdir -p gpurun_out
bash tools/dev_variants.sh "$@" 2>&1 | tee gpurun_out/variants.txt
bash tools/dev_variants.sh "$@" -- --holes 0 2>&1 | tee gpurun_out/variants_noholes.txt
python tools/dev_scale_check.py 2>&1 | tail -6 | tee gpurun_out/scale_base.txt
if [ -n "$1" ]; then
  TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_$1.so python tools/dev_scale_check.py 2>&1 | tail -6 | tee gpurun_out/scale_$1.txt
  TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_$1.so python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests_$1.txt
fi
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests_base.txt
