set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python scripts/diag_mcc.py 48 smooth | cut -c1-120
python bench.py > gpurun_out/bench_n1_uniform.json 2> gpurun_out/bench_n1_uniform.err; tail -c 2600 gpurun_out/bench_n1_uniform.json
python bench.py --kind smooth --no-cpu-baseline > gpurun_out/bench_n1_smooth.json 2>/dev/null; tail -c 1700 gpurun_out/bench_n1_smooth.json
