#!/bin/bash
# ncu launch list of eager iterations of the bench workload (per-launch gpu__time_duration; cold-cache, serialised: compare shares)
# usage (under gpurun): bash tools/launch_list.sh gpurun_out/launches.csv
out=${1:-gpurun_out/launches.csv}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out" python bench.py --eager --steps 2 --warmup 3 > gpurun_out/launch_list_bench.log 2>&1
python - "$out" <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki][:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("kernels launched: %d, summed duration %.1f us" % (sum(a[0] for a in agg.values()), tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6d  %9.1f us  %5.1f%%  %s" % (a[0], a[1] / 1e3, 100 * a[1] / tot, k))
PY
