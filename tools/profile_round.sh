#!/bin/bash
# Round-2 evidence run (one B200): plain bench, ncu launch list of the same command, ncu --set full of the hot kernels.
# The .ncu-rep is summarised on the box and removed (gpurun brings back <= 64 MiB).
set -x
export EAVIT_STEP_GRAPH=0     # eager launches: one ncu row per kernel launch, attributable CUDA-event table
python bench.py --steps 2 --warmup 1 --no-side --no-cpu --no-e2e > gpurun_out/r2_plain_for_ncu.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 460 --csv --log-file gpurun_out/r2_ncu_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --no-side --no-cpu --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
python tools/ncu_launch_shares.py gpurun_out/r2_ncu_launch_list.csv > gpurun_out/r2_ncu_launch_shares.md
python tools/prof_kernels.py all > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'attention|gemm_bf16|layernorm|rms_u8x16|obs_normalize|rms_partial|rms_reduce|embed_fused' -c 56 \
    -o /tmp/r2_prof python tools/prof_kernels.py all > gpurun_out/r2_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r2_prof.ncu-rep > gpurun_out/r2_ncu_full_top_kernels_table.md
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv 2>/dev/null | cut -c1-100000 | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keep=[i for i,c in enumerate(h) if c in ('Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] if i<len(r) else '' for i in keep])
" > gpurun_out/r2_ncu_full_raw_selected.csv
ls -la gpurun_out/r2_*
