set -x
O=gpurun_out/r02_call6; mkdir -p $O
for v in 0 1 2; do ./scripts/probes/tma_probe $v 256; ./scripts/probes/tma_probe $v 64; done 2>&1 | tee $O/tma_probe.txt
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_matrices_gpu.py -q -m gpu -x -k "tma_equals" 2>&1 | grep -v "^$" | head -60 | cut -c1-250 | tee $O/tma_memcheck.txt
B200_SEG_TMA=0 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 | tee $O/pytest_notma.txt
B200_SEG_TMA=0 python scripts/prof_segment.py 256 smooth 3 2>&1 | tee $O/segment_wall_notma.txt
B200_SEG_TMA=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/segment_launches.csv python scripts/prof_segment.py 256 smooth 1 > /dev/null 2>&1
python - "$O/segment_launches.csv" <<'PY' | tee $O/segment_launches.txt
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0][-50:]
    v = float(r[iv].replace(",", "")); v = v / 1e6 if r[iu] == "ns" else v / 1e3 if r[iu] == "us" else v
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print(f"{k:52s} n={len(v)} total {sum(v):8.3f} ms")
PY
