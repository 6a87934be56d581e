set -x
O=gpurun_out/r02_call19; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/d2h_probe.py > $O/d2h_probe_n8.json 2> $O/d2h_probe.err
cat $O/d2h_probe_n8.json; tail -3 $O/d2h_probe.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 1500 $O/bench_n8.json; tail -3 $O/bench_n8.err
