#!/bin/bash
# Round-2 profile captures (run under gpurun, one GPU): every command runs plainly first, then under ncu.
#   1. launch list of the bench command (device time of every kernel of 2 timed sweeps of the full C4 config)
#   2. --set full capture of the sampling kernel's class launches on a 500 k-document C4 slice
#   3. DRAM traffic + instruction counts of the sampling kernel for C2, C3 and the full C4 (sweep 2)
out=gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --after-sweeps 0"
$B --steps 2 --warmup 1 > $out/r02_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_c4.csv $B --steps 2 --warmup 1 > $out/r02_ncu_launches.log 2>&1
$B --docs 500000 --steps 2 --warmup 1 > $out/r02_plain_c4s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gibbs_sweep -s 5 -c 5 -o $out/r02_sweep_final $B --docs 500000 --steps 2 --warmup 1 > $out/r02_ncu_full.log 2>&1
for wl in c2 c3 c4; do
  $B --workload $wl --steps 1 --warmup 1 > $out/r02_plain_$wl.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:k_gibbs_sweep -c 40 --csv --log-file $out/r02_traffic_$wl.csv $B --workload $wl --steps 1 --warmup 1 > $out/r02_ncu_traffic_$wl.log 2>&1
done
tail -1 $out/r02_plain_c4.log | cut -c1-300
