#!/bin/bash
# Round-2 evidence run (one B200): plain bench, ncu launch list of the same command, ncu --set full of the hot kernels.
set -x
export EAVIT_STEP_GRAPH=0     # eager launches: one ncu row per kernel launch, attributable CUDA-event table
python bench.py --steps 2 --warmup 1 --no-side --no-cpu --no-e2e > gpurun_out/r2_plain_for_ncu.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 460 --csv --log-file gpurun_out/r2_ncu_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --no-side --no-cpu --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
python tools/prof_kernels.py all > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attention|gemm_bf16|layernorm|rms_u8x16|obs_normalize|rms_partial' -c 60 \
    -o gpurun_out/r2_prof python tools/prof_kernels.py all > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_prof* gpurun_out/r2_ncu*
