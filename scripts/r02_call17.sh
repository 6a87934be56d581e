set -x
O=gpurun_out/r02_call17; mkdir -p $O
{
for nt in 512 640 768; do echo "B200_GLCM_NT=$nt"; for k in uniform smooth; do B200_GLCM_NT=$nt python scripts/quick_time.py 256 $k | grep -E "glcm|glszm"; done; done
} 2>&1 | grep -v "^+" | tee $O/quick_time_nt.txt
timeout 1200 python -m pytest tests/test_voxel_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_voxel.txt
B200_GLCM_NT=640 python scripts/diag_mcc.py 64 2>&1 | cut -c1-160 | tee $O/diag_nt640.txt
