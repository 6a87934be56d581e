set -x
O=gpurun_out/r02_call18; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
nvidia-smi dmon -s put -d 1 -o T > $O/dmon.txt 2>&1 &
DM=$!
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
kill $DM
tail -c 2500 $O/bench_n8.json; tail -5 $O/bench_n8.err
