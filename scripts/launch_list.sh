# per-kernel durations (ncu, single metric) of one GLCM call: scripts/launch_list.sh SIZE KIND OUT.csv
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:glcm_fast --launch-skip ${4:-25} -c ${4:-25} --csv --log-file $3 python scripts/prof_glcm.py $1 $2 glcm > /dev/null 2>&1
python - "$3" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0][-40:]
    v = float(r[iv].replace(",", "")); v = v / 1e6 if r[iu] == "ns" else v / 1e3 if r[iu] == "us" else v
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print(f"{k:42s} n={len(v)} total {sum(v):8.2f} ms  first {v[0]:7.3f} ms")
PY
