#!/bin/bash
# Round-2 profile set (run under gpurun): launch list of a short bench run + full captures of the two hot kernels.
set -x
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
python tools/c5_only.py > gpurun_out/r2_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_r2_c5_final -f \
    python tools/c5_only.py > gpurun_out/r2_ncu_c5.log 2>&1
python tools/c5_only.py canonical > gpurun_out/r2_c5c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_r2_c5_canonical -f \
    python tools/c5_only.py canonical > gpurun_out/r2_ncu_c5c.log 2>&1
python tools/c4_only.py > gpurun_out/r2_c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_small -s 4 -c 1 -o gpurun_out/prof_r2_c4_final -f \
    python tools/c4_only.py > gpurun_out/r2_ncu_c4.log 2>&1
# summaries are made on the box (gpurun brings back at most 64 MiB): text files only, the reports are dropped
python tools/ncu_summary.py gpurun_out/prof_r2_c5_final.ncu-rep > gpurun_out/r2_c5_final_ncu_summary.txt
python tools/ncu_lines.py gpurun_out/prof_r2_c5_final.ncu-rep 60 > gpurun_out/r2_c5_final_ncu_lines.txt
python tools/ncu_summary.py gpurun_out/prof_r2_c5_canonical.ncu-rep > gpurun_out/r2_c5_canonical_ncu_summary.txt
python tools/ncu_summary.py gpurun_out/prof_r2_c4_final.ncu-rep > gpurun_out/r2_c4_pixel_ncu_summary.txt
python tools/ncu_lines.py gpurun_out/prof_r2_c4_final.ncu-rep 60 > gpurun_out/r2_c4_pixel_ncu_lines.txt
ncu -i gpurun_out/prof_r2_c5_final.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/r2_c5_final_traffic.csv
rm -f gpurun_out/prof_r2_c5_canonical.ncu-rep gpurun_out/prof_r2_c4_final.ncu-rep gpurun_out/prof_r2_c5_final.ncu-rep
for f in gpurun_out/r2_ncu_c5.log gpurun_out/r2_ncu_c5c.log gpurun_out/r2_ncu_c4.log; do tail -n 2 $f; done
