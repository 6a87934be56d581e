set -x
O=gpurun_out/r02_call9; mkdir -p $O
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 | tee $O/pytest.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_all.py 16 matrix filters shape firstorder > $O/memcheck.log 2>&1; echo rc=$? >> $O/memcheck.log; tail -3 $O/memcheck.log
timeout 300 compute-sanitizer --tool racecheck python scripts/sanitize_all.py 16 matrix filters > $O/racecheck.log 2>&1; echo rc=$? >> $O/racecheck.log; tail -3 $O/racecheck.log
python scripts/prof_segment.py 256 smooth 3 2>&1 | grep -v "^+" | tee $O/segment_wall_tma.txt
B200_SEG_TMA=0 python scripts/prof_segment.py 256 smooth 3 2>&1 | grep -v "^+" | tee $O/segment_wall_coop.txt
for m in 1 0; do B200_SEG_TMA=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/segment_launches_tma$m.csv python scripts/prof_segment.py 256 smooth 1 > /dev/null 2>&1; done
python - $O/segment_launches_tma1.csv $O/segment_launches_tma0.csv <<'PY' | tee $O/segment_launches.txt
import csv, sys, collections
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0][-50:]
        v = float(r[iv].replace(",", "")); v = v / 1e6 if r[iu] == "ns" else v / 1e3 if r[iu] == "us" else v
        agg.setdefault(name, []).append(v)
    print(f)
    for k, v in agg.items(): print(f"  {k:52s} n={len(v)} total {sum(v):8.3f} ms")
PY
timeout 300 ncu --set full --clock-control none -k regex:"seg_tile|seg_glrlm_ends|swt3d|recursive_gauss_x" -c 6 -o $O/ncu_segment python scripts/sanitize_all.py 64 matrix filters > /dev/null 2>&1
(time python bench.py --steps 5 --warmup 3) > $O/bench_512.json 2> $O/bench_512.err; tail -3 $O/bench_512.err; cut -c1-300 $O/bench_512.json
