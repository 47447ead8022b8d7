#!/bin/bash
# Device time of the sampling kernel (sweeps 1-20, C4 / C2 / C3 shapes) for several builds of the library.
#   tools/variant_sweep.sh <lib.so> [<lib.so> ...]     ("default" = the in-tree build)
for lib in "$@"; do
  for wl in ${WORKLOADS:-c4 c2 c3}; do
    docs=$([ $wl = c4 ] && echo 1000000 || echo 300000)
    if [ "$lib" = default ]; then unset B200LDA_LIB; else export B200LDA_LIB=$PWD/$lib; fi
    timeout 300 python tools/sweep_trajectory.py --workload $wl --docs $docs --mode ${MODE:-live} --sweeps 20 --every 20 2>&1 | tail -1 |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$wl', round(d['sample_ms'],2),'ms', round(d['tok_per_s']/1e9,3),'Gtok/s', 'kd',round(d['kd'],1),'moved',round(d['moved'],2),'prior',round(d['prior'],2))"
  done
done
