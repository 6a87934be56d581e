# round 2, first GPU call: compute-sanitizer over every kernel family, the GF_LZ_SMEM diagnosis, environment probes,
# and the starting-point timings.  gpurun --timeout 1500 -- 'bash scripts/r02_call1.sh'
set -x
O=gpurun_out/r02_call1; mkdir -p $O
python -c "import pywt; print('pywt', pywt.__version__)" 2>&1 | tail -1 | tee $O/probe_libs.txt
python -c "import SimpleITK as s; print('sitk', s.__version__)" 2>&1 | tail -1 | tee -a $O/probe_libs.txt
(nvidia-smi topo -m; lscpu | grep -E "NUMA|Model name|^CPU\(s\)|Socket"; nproc) > $O/topo.txt 2>&1
timeout 420 compute-sanitizer --tool memcheck python scripts/sanitize_all.py 16 > $O/memcheck.log 2>&1; echo rc=$? >> $O/memcheck.log; tail -4 $O/memcheck.log
timeout 480 compute-sanitizer --tool racecheck python scripts/sanitize_all.py 12 fast matrix filters shape firstorder > $O/racecheck.log 2>&1; echo rc=$? >> $O/racecheck.log; tail -4 $O/racecheck.log
timeout 200 compute-sanitizer --tool synccheck python scripts/sanitize_all.py 12 fast generic > $O/synccheck.log 2>&1; echo rc=$? >> $O/synccheck.log; tail -3 $O/synccheck.log
V=pyradiomics_b200/variants
B200_RADIOMICS_LIB=$V/libsmem166.so python scripts/diag_mcc.py 48 smooth 2>&1 | cut -c1-200 > $O/smem166_diag.log; tail -3 $O/smem166_diag.log
B200_RADIOMICS_LIB=$V/libsmem166.so timeout 240 compute-sanitizer --tool racecheck python scripts/diag_mcc.py 20 smooth > $O/smem166_racecheck.log 2>&1; tail -5 $O/smem166_racecheck.log | cut -c1-200
B200_RADIOMICS_LIB=$V/libsmem166.so timeout 200 compute-sanitizer --tool memcheck python scripts/diag_mcc.py 20 smooth > $O/smem166_memcheck.log 2>&1; tail -5 $O/smem166_memcheck.log | cut -c1-200
for k in uniform smooth; do python scripts/quick_time.py 256 $k; done 2>&1 | tee $O/quick_time_256.txt
