# the last GPU call of round 2: full GPU test suite, smoke(), the driver's bench command, ncu evidence for profiles/
set -x
O=gpurun_out/r02_final; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 1500 $O/bench_n1.json; tail -3 $O/bench_n1.err
# launch list of the bench command at 256^3 (shares must agree with the live CUDA-event numbers)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench256.csv python bench.py --size 256 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $O/bench256_under_ncu.log 2>&1
python - $O/launches_bench256.csv > $O/launches_bench256.md <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0][-48:]
    v = float(r[iv].replace(",", "")); v = v / 1e6 if r[iu] == "ns" else v / 1e3 if r[iu] == "us" else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])): print(f"| `{k}` | {len(v)} | {sum(v):.2f} | {100*sum(v)/tot:.1f} % |")
PY
cat $O/launches_bench256.md
gzip -9 $O/launches_bench256.csv
# ncu --set full of one plane chunk of the GLCM kernels (8 launches with per-group solves), both volumes
R=/tmp/reps; mkdir -p $R
for k in uniform smooth; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:glcm_fast -s 40 -c 8 -o $R/ncu_glcm_256_$k python scripts/prof_glcm.py 256 $k glcm > /dev/null 2>&1
  python scripts/summarize_ncu.py $R/ncu_glcm_256_$k.ncu-rep > $O/ncu_glcm_256_$k.txt 2>&1
  ncu -i $R/ncu_glcm_256_$k.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size > $O/ncu_glcm_256_$k.raw.csv 2>&1
done
for c in glrlm glszm; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:"glrlm_fast_kernel|small_fast_kernel" -s 1 -c 1 -o $R/ncu_$c python scripts/prof_glcm.py 256 uniform $c > /dev/null 2>&1
  python scripts/summarize_ncu.py $R/ncu_$c.ncu-rep > $O/ncu_${c}_256_uniform.txt 2>&1
done
ls -la $O $R
