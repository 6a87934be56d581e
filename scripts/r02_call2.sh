# round 2, call 2: the register Lanczos solver + phase-A bipartite filter on hardware: parity, determinism, time, ncu
set -x
O=gpurun_out/r02_call2; mkdir -p $O
python -m pytest tests/test_voxel_gpu.py tests/test_plugins_gpu.py -q -m gpu -x 2>&1 | tail -5 | tee $O/pytest.txt
python scripts/diag_mcc.py 64 2>&1 | cut -c1-220 | tee $O/diag_mcc.txt
for k in uniform smooth; do python scripts/quick_time.py 256 $k; done 2>&1 | grep -v "^+" | tee $O/quick_time_256.txt
bash scripts/launch_list.sh 256 uniform $O/launches_uniform.csv 2>&1 | tail -12 | tee $O/launches_uniform.txt
bash scripts/launch_list.sh 256 smooth $O/launches_smooth.csv 2>&1 | tail -12 | tee $O/launches_smooth.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glcm_fast_solve_kernel -s 3 -c 3 -o $O/ncu_solve_uniform python scripts/prof_glcm.py 160 uniform glcm > $O/ncu_solve.log 2>&1; tail -2 $O/ncu_solve.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glcm_fast_kernel -s 1 -c 1 -o $O/ncu_phaseA_uniform python scripts/prof_glcm.py 160 uniform glcm > $O/ncu_phaseA.log 2>&1; tail -2 $O/ncu_phaseA.log
ls -la $O
