set -x
python bench.py --kind smooth --no-cpu-baseline > gpurun_out/bench_n1_smooth.json 2>/dev/null; tail -c 1700 gpurun_out/bench_n1_smooth.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench256.csv python bench.py --size 256 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --set full --clock-control none -k regex:'glcm_fast' --launch-skip 25 -c 5 -o /tmp/r01_glcm_full -f python scripts/prof_glcm.py 256 uniform glcm > gpurun_out/ncu_glcm.log 2>&1
python scripts/summarize_ncu.py /tmp/r01_glcm_full.ncu-rep > gpurun_out/r01_ncu_glcm_256_uniform.txt 2>&1
ncu --set full --clock-control none -k regex:'glcm_fast' --launch-skip 25 -c 5 -o /tmp/r01_glcm_smooth_full -f python scripts/prof_glcm.py 256 smooth glcm > gpurun_out/ncu_glcm_s.log 2>&1
python scripts/summarize_ncu.py /tmp/r01_glcm_smooth_full.ncu-rep > gpurun_out/r01_ncu_glcm_256_smooth.txt 2>&1
cp /tmp/r01_glcm_full.ncu-rep gpurun_out/
ls -la gpurun_out
