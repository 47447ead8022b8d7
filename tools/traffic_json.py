#!/usr/bin/env python
"""profiles/<round>_traffic_<workload>.json from the ncu metric list tools/profile_r02*.sh writes.
usage: traffic_json.py <ncu.csv> <plain bench log> <workload> <sweep> > profiles/r02_traffic_c4.json
The capture holds every sampling-kernel launch of the run (warm-up sweep first); <sweep> (1-based)
selects which sweep's class launches to sum."""
import csv, json, sys

csv_path, plain_log, workload, sweep = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
h = rows[0]
col = {n: h.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value")}
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[col["ID"]]), {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]]})
    d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
bench = json.loads(open(plain_log).read().strip().splitlines()[-1])
sweeps_in_run = bench["steps"] + bench["warmup"]
per_sweep = len(launches) // sweeps_in_run
ids = sorted(launches)[(sweep - 1) * per_sweep: sweep * per_sweep]
tokens = bench["config"]["tokens"]
cls = [{"kernel": launches[i]["kernel"], "grid": launches[i]["grid"],
        "ms": launches[i]["gpu__time_duration.sum"] / 1e6,
        "dram_read": launches[i]["dram__bytes_read.sum"], "dram_write": launches[i]["dram__bytes_write.sum"],
        "inst": launches[i]["smsp__inst_executed.sum"], "l2_hit_pct": launches[i]["lts__t_sector_hit_rate.pct"],
        "issue_active_pct": launches[i]["smsp__issue_active.avg.pct_of_peak_sustained_active"]} for i in ids]
dram = sum(c["dram_read"] + c["dram_write"] for c in cls)
print(json.dumps({
    "workload": workload, "tokens": tokens, "K": bench["config"].get("K"), "sweep": sweep, "launches": len(cls),
    "traffic_bytes_per_token": dram / tokens, "dram_bytes": dram,
    "warp_instructions_per_token": sum(c["inst"] for c in cls) / tokens,
    "kernel_ms_serialised_by_ncu": sum(c["ms"] for c in cls), "kernel_ms_plain_run": bench["ms_per_step"],
    "classes": cls,
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,... "
              "--clock-control none -k regex:k_gibbs_sweep -c 40 ; python bench.py --workload %s --steps 1 --warmup 1 "
              "--no-cpu-baseline --e2e-steps 0 --after-sweeps 0 (sweep %d: its row-width class launches, serialised by ncu; "
              "tools/profile_r02_final.sh, tools/traffic_json.py)" % (workload, sweep)}, indent=1))
