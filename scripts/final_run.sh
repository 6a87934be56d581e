set -x
python bench.py > gpurun_out/bench_n1_uniform.json 2> gpurun_out/bench_n1_uniform.err; tail -c 3000 gpurun_out/bench_n1_uniform.json
python bench.py --kind smooth --no-cpu-baseline > gpurun_out/bench_n1_smooth.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench256.csv python bench.py --size 256 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b.log 2>&1
for c in glcm ngtdm gldm; do ncu --set full --clock-control none --import-source on -k regex:'glcm_fast|small_fast' --launch-skip $([ $c = glcm ] && echo 15 || echo 1) -c $([ $c = glcm ] && echo 3 || echo 1) -o gpurun_out/r01_${c}_full -f python scripts/prof_glcm.py 256 uniform $c > gpurun_out/ncu_$c.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:'glcm_fast_solve' --launch-skip 5 -c 1 -o gpurun_out/r01_glcm_solve_smooth_full -f python scripts/prof_glcm.py 256 smooth glcm > gpurun_out/ncu_solve.log 2>&1
ls -la gpurun_out
