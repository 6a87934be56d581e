# First GPU call of the next round: the experiments prepared (CPU-verified, not yet run on hardware) at the end of
# round 1, in ONE box session.  Build the variants on the CPU box first:
#   python -m pyradiomics_b200.build --variant smem64 GF_LZ_SMEM=1 GF_SOLVE_MINB_L=8
#   python -m pyradiomics_b200.build --variant smem166 GF_LZ_SMEM=1 GF_SOLVE_MINB_L=3
#   python -m pyradiomics_b200.build --variant tile4k GF_SOLVE_TILE=4096
# then: gpurun --timeout 900 -- 'bash scripts/next_round_first_call.sh 2>&1 | tail -60'
set -x
V=pyradiomics_b200/variants
# 1. is the shared-memory Lanczos scratch a register-allocation or a memory problem?  (determinism + parity probe)
for v in smem64 smem166; do B200_RADIOMICS_LIB=$V/lib$v.so python scripts/diag_mcc.py 48 smooth | cut -c1-140; done
# 1b. racecheck / memcheck of the failing variant on a small volume (slow: keep it tiny)
B200_RADIOMICS_LIB=$V/libsmem166.so timeout 300 compute-sanitizer --tool racecheck python scripts/diag_mcc.py 24 smooth 2>&1 | tail -15
B200_RADIOMICS_LIB=$V/libsmem166.so timeout 300 compute-sanitizer --tool memcheck python scripts/diag_mcc.py 24 smooth 2>&1 | tail -15
# 2. two-stream overlap of phase A and the eigen-solves: correctness first, then time
B200_GLCM_OVERLAP=1 python scripts/diag_mcc.py 48 | cut -c1-140
B200_GLCM_OVERLAP=1 python -m pytest tests/test_voxel_gpu.py -q -m gpu -x 2>&1 | tail -2
for k in uniform smooth; do
  python scripts/quick_time.py 256 $k | grep -E "glcm|suite"
  B200_GLCM_OVERLAP=1 python scripts/quick_time.py 256 $k | grep -E "glcm|suite"
  B200_RADIOMICS_LIB=$V/libtile4k.so python scripts/quick_time.py 256 $k | grep -E "glcm|suite"
done
