set -x
O=gpurun_out/r02_call3; mkdir -p $O
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee $O/pytest.txt
python scripts/diag_mcc.py 64 2>&1 | cut -c1-200 | tee $O/diag_mcc.txt
for nt in 256 384 512; do for k in uniform smooth; do echo "NT=$nt"; B200_GLCM_NT=$nt python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done 2>&1 | grep -v "^+" | tee $O/quick_time_nt.txt
python scripts/quick_time.py 256 smooth 2>&1 | tee $O/quick_time_smooth.txt
bash scripts/launch_list.sh 256 smooth $O/launches_smooth.csv 2>&1 | tail -8 | tee $O/launches_smooth.txt
(time python bench.py --steps 5 --warmup 3) > $O/bench_512.json 2> $O/bench_512.err; tail -5 $O/bench_512.err; cut -c1-1500 $O/bench_512.json
