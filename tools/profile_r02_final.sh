#!/bin/bash
# Re-capture for the final round-2 kernel (per-lane dead masks), one GPU, every command plainly first:
#   1. launch list of 2 timed sweeps of the full C4 config (only this library's kernels)
#   2. --set full capture of the sampling kernel's class launches on a 500 k-document C4 slice
#   3. DRAM traffic + instruction counts of the sampling kernel on the full C4 (sweep 2)
out=gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 --after-sweeps 0"
$B --steps 2 --warmup 1 > $out/r02f_plain_c4.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $out/r02f_launches_c4.csv $B --steps 2 --warmup 1 > $out/r02f_ncu_launches.log 2>&1
$B --docs 500000 --steps 2 --warmup 1 > $out/r02f_plain_c4s.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gibbs_sweep -s 5 -c 5 -f -o $out/r02f_sweep_final $B --docs 500000 --steps 2 --warmup 1 > $out/r02f_ncu_full.log 2>&1
$B --steps 1 --warmup 1 > $out/r02f_plain_c4b.log 2>&1 &&
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_gibbs_sweep -c 40 --csv --log-file $out/r02f_traffic_c4.csv $B --steps 1 --warmup 1 > $out/r02f_ncu_traffic_c4.log 2>&1
tail -1 $out/r02f_plain_c4.log | cut -c1-300
ls -la $out | grep r02f
