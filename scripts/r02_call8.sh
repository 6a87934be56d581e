set -x
O=gpurun_out/r02_call8; mkdir -p $O
P=./scripts/probes/tma_probe
( $P 0 256 64 4 2 0 0 0; $P 0 256 80 10 6 0 0 0; $P 0 256 64 4 2 -1 -1 7; $P 0 256 80 4 2 0 0 0; $P 0 256 64 10 6 0 0 0; $P 0 256 80 10 6 16 1 1; $P 0 256 96 10 6 -16 -1 7; $P 0 256 80 10 6 -1 0 0 ) 2>&1 | grep -v "^+" | tee $O/tma_probe3.txt
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 | tee $O/pytest.txt
