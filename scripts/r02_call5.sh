set -x
O=gpurun_out/r02_call5; mkdir -p $O
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-400 | tee $O/pytest.txt
timeout 420 compute-sanitizer --tool memcheck python scripts/sanitize_all.py 16 > $O/memcheck.log 2>&1; echo rc=$? >> $O/memcheck.log; tail -3 $O/memcheck.log
timeout 480 compute-sanitizer --tool racecheck python scripts/sanitize_all.py 12 fast matrix filters > $O/racecheck.log 2>&1; echo rc=$? >> $O/racecheck.log; tail -3 $O/racecheck.log
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
B200_SEG_TMA=0 python scripts/prof_segment.py 256 smooth 2 2>&1 | grep -v "^+" | tee $O/segment_wall_notma.txt
for v in "" glrlm3 glrlm2; do echo "variant=$v"; if [ -n "$v" ]; then export B200_RADIOMICS_LIB=pyradiomics_b200/variants/lib$v.so; fi; for k in uniform smooth; do python scripts/quick_time.py 256 $k | grep -E "glrlm|ngtdm|glcm"; done; done 2>&1 | grep -v "^+" | tee $O/quick_time_variants.txt
unset B200_RADIOMICS_LIB
bash scripts/launch_list.sh 256 uniform $O/launches_uniform.csv 2>&1 | tail -6 | tee $O/launches_uniform.txt
bash scripts/launch_list.sh 256 smooth $O/launches_smooth.csv 2>&1 | tail -6 | tee $O/launches_smooth.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:glrlm_fast_kernel -s 1 -c 1 -o $O/ncu_glrlm python scripts/prof_glcm.py 160 uniform glrlm > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:glcm_fast_kernel -s 1 -c 1 -o $O/ncu_phaseA512 python scripts/prof_glcm.py 160 uniform glcm > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"seg_tile|seg_glrlm_ends|ccl_merge" -c 3 -o $O/ncu_segment python scripts/prof_segment.py 256 smooth 1 > /dev/null 2>&1
(time python bench.py --size 256 --steps 3 --warmup 3 --parity-voxels 2048 --no-cpu-baseline) > $O/bench_256.json 2> $O/bench_256.err; tail -3 $O/bench_256.err
ls -la $O
