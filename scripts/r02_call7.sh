set -x
O=gpurun_out/r02_call7; mkdir -p $O
for t in 1 2 3 4 5; do ./scripts/probes/tma_probe2 $t; done 2>&1 | grep -v "^+" | tee $O/tma_probe2.txt
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 | tee $O/pytest.txt
python scripts/prof_segment.py 256 uniform 3 2>&1 | grep -v "^+" | tee $O/segment_wall.txt
for k in uniform smooth; do python scripts/quick_time.py 256 $k; done 2>&1 | grep -v "^+" | tee $O/quick_time_256.txt
(time python bench.py --size 256 --steps 3 --warmup 3 --parity-voxels 2048) > $O/bench_256.json 2> $O/bench_256.err; tail -3 $O/bench_256.err
