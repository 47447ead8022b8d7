#!/bin/bash
# The measurement behind every row of profiles/r01_tuning.md: device time of the sampling kernel per
# sweep (sweeps 1-20 from random init) on the C4 / C2 / C3 shapes. Run from the repo root on a B200.
#   tools/quick_sweep.sh [live|deferred]
mode=${1:-live}
for wl in c4 c2 c3; do
  docs=$([ $wl = c4 ] && echo 1000000 || echo 300000)
  timeout 300 python tools/sweep_trajectory.py --workload $wl --docs $docs --mode $mode --sweeps 20 --every 20 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', '$mode', round(d['sample_ms'],2),'ms', round(d['tok_per_s']/1e9,3),'Gtok/s', 'kd',round(d['kd'],1),'moved',round(d['moved'],2),'prior',round(d['prior'],2))"
done
