set -x
O=gpurun_out/r02_call16; mkdir -p $O
for k in uniform smooth; do python scripts/quick_time.py 256 $k; done 2>&1 | grep -v "^+" | tee $O/quick_time.txt
timeout 1200 python -m pytest tests/test_voxel_gpu.py tests/test_plugins_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee $O/pytest_voxel_plugins.txt
python scripts/prof_host.py 256 256 2>&1 | grep -E "suite" | tee $O/prof_host_suite.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 6000 $O/bench_n1.json
tail -5 $O/bench_n1.err
