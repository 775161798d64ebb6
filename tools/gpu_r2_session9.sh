#!/bin/bash
# Round-2 session 9: the MATH2 variant of the fused kernel (A/B + whole-map check + suite), then the suite on the product library.
mkdir -p gpurun_out
bash tools/dev_variants.sh m2 2>&1 | tee gpurun_out/variants9.txt
bash tools/dev_variants.sh m2 -- --holes 0 2>&1 | tee gpurun_out/variants9_noholes.txt
TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_m2.so python tools/dev_scale_check.py 2>&1 | tail -6 | tee gpurun_out/scale9_m2.txt
TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_m2.so python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests9_m2.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests9_base.txt
