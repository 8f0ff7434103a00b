#!/bin/bash
# DRAM traffic + duration of the fused C5 kernel for several builds (exploration helper; run under gpurun)
for l in "" $@; do
  echo "== lib=$l"
  RRT_B200_LIB=$l python tools/c5_time.py 2>&1 | grep "sweep=0"
  RRT_B200_LIB=$l ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:render_kernel -s 3 -c 1 python tools/c5_only.py 2>&1 | grep -E "dram__|gpu__time|l1tex"
done
