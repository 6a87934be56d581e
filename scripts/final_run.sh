# round artefacts on one box: bench lines, launch list, ncu full capture of one GLCM plane chunk
set -x
python bench.py > gpurun_out/bench_n1_uniform.json 2> gpurun_out/bench_n1_uniform.err; tail -c 2500 gpurun_out/bench_n1_uniform.json
python bench.py --kind smooth --no-cpu-baseline > gpurun_out/bench_n1_smooth.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench256.csv python bench.py --size 256 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --set full --clock-control none -k regex:'glcm_fast' --launch-skip 25 -c 5 -o gpurun_out/r01_glcm_full -f python scripts/prof_glcm.py 256 uniform glcm > gpurun_out/ncu_glcm.log 2>&1
ncu --set full --clock-control none -k regex:'glcm_fast' --launch-skip 25 -c 5 -o gpurun_out/r01_glcm_smooth_full -f python scripts/prof_glcm.py 256 smooth glcm > gpurun_out/ncu_glcm_s.log 2>&1
ls -la gpurun_out
