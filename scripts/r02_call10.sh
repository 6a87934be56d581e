set -x
O=gpurun_out/r02_call10; mkdir -p $O
python -m pytest tests/test_multigpu_filters.py tests/test_resample_gpu.py -q -m gpu -x 2>&1 | tail -6 | cut -c1-400 | tee $O/pytest_2gpu.txt
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > $O/bench_n2.json 2> $O/bench_n2.err; tail -5 $O/bench_n2.err | cut -c1-300; cut -c1-400 $O/bench_n2.json
