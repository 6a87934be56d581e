set -x
O=gpurun_out/r02_call4; mkdir -p $O
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 | tee $O/pytest.txt
python scripts/diag_mcc.py 64 2>&1 | cut -c1-160 | tee $O/diag_mcc.txt
for nt in 256 384 512; do for k in uniform smooth; do echo "NT=$nt"; B200_GLCM_NT=$nt python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done 2>&1 | grep -v "^+" | tee $O/quick_time_nt.txt
for k in uniform smooth; do python scripts/quick_time.py 256 $k; done 2>&1 | grep -v "^+" | tee $O/quick_time_256.txt
for k in uniform smooth; do python scripts/prof_segment.py 256 $k 2; done 2>&1 | grep -v "^+" | tee $O/segment_wall.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/segment_launches.csv python scripts/prof_segment.py 256 smooth 1 > /dev/null 2>&1
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
(time python bench.py --size 256 --steps 3 --warmup 3 --parity-voxels 4096) > $O/bench_256.json 2> $O/bench_256.err; tail -4 $O/bench_256.err; cut -c1-600 $O/bench_256.json
