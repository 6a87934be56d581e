set -x
O=gpurun_out/r02_call13; mkdir -p $O
V=pyradiomics_b200/variants
for v in "" eigstat2; do echo "variant=$v"; if [ -n "$v" ]; then export B200_RADIOMICS_LIB=$V/lib$v.so; else unset B200_RADIOMICS_LIB; fi; for k in uniform smooth; do python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done 2>&1 | grep -v "^+" | tee $O/quick_time_variants.txt
B200_RADIOMICS_LIB=$V/libeigstat2.so python scripts/diag_mcc.py 96 2>&1 | cut -c1-160 | tee $O/diag_eigstat2.txt
for v in "" eigstat2; do
  if [ -n "$v" ]; then export B200_RADIOMICS_LIB=$V/lib$v.so; else unset B200_RADIOMICS_LIB; fi
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:glcm_fast_solve_kernel -s 17 -c 1 -o $O/ncu_lanczos_256_uniform_$v python scripts/prof_glcm.py 256 uniform glcm > $O/ncu_lz_$v.log 2>&1
  python scripts/summarize_ncu.py $O/ncu_lanczos_256_uniform_$v.ncu-rep > $O/ncu_lanczos_256_uniform_$v.txt 2>&1
  ncu -i $O/ncu_lanczos_256_uniform_$v.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$v.csv 2>$O/src_$v.err
  python scripts/src_lines.py /tmp/src_$v.csv 70 > $O/lines_lanczos_$v.txt 2>&1
done
unset B200_RADIOMICS_LIB
ls -la $O
