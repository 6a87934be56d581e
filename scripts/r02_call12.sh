set -x
O=gpurun_out/r02_call12; mkdir -p $O
V=pyradiomics_b200/variants
for v in "" segsum0 tile8k tile2k; do echo "variant=$v"; if [ -n "$v" ]; then export B200_RADIOMICS_LIB=$V/lib$v.so; else unset B200_RADIOMICS_LIB; fi; for k in uniform smooth; do python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done 2>&1 | grep -v "^+" | tee $O/quick_time_variants.txt
unset B200_RADIOMICS_LIB
python scripts/diag_mcc.py 96 2>&1 | cut -c1-160 | tee $O/diag.txt
R=/tmp/reps; mkdir -p $R
summ() { # rep name kernels-for-source...
  python scripts/summarize_ncu.py $R/$1.ncu-rep > $O/$1.txt 2>&1
  ncu -i $R/$1.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size > $O/$1.raw.csv 2>&1
}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:glcm_fast -s 25 -c 5 -o $R/ncu_glcm_256_uniform python scripts/prof_glcm.py 256 uniform glcm > /dev/null 2>&1
summ ncu_glcm_256_uniform
timeout 600 ncu --set full --import-source on --clock-control none -k regex:glcm_fast -s 25 -c 5 -o $R/ncu_glcm_256_smooth python scripts/prof_glcm.py 256 smooth glcm > /dev/null 2>&1
summ ncu_glcm_256_smooth
# source-line shares of the Lanczos solve and phase A (one launch each)
for k in "glcm_fast_solve_kernel<2>" "glcm_fast_kernel"; do
  tag=$(echo $k | tr -d '<>' ); 
  ncu -i $R/ncu_glcm_256_uniform.ncu-rep --page source --csv --print-source cuda,sass -k regex:"$(echo $k | sed 's/[<>]/./g')" > /tmp/src_$tag.csv 2>/dev/null
  python scripts/src_lines.py /tmp/src_$tag.csv 60 > $O/lines_uniform_$tag.txt 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"seg_tile|seg_glrlm_ends|ccl_merge" -c 4 -o $R/ncu_segment_256 python scripts/prof_segment.py 256 smooth 1 > /dev/null 2>&1
summ ncu_segment_256
timeout 300 ncu --set full --clock-control none -k regex:"swt3d|recursive_gauss" -c 4 -o $R/ncu_filters_256 python - <<'PY' > /dev/null 2>&1
import sys, torch
sys.path.insert(0, ".")
from pyradiomics_b200 import imageoperations as IO
x = torch.randn((256, 256, 256), device="cuda", dtype=torch.float64)
lo, hi = IO.wavelet_filters("coif1")
IO.swt_level1_device(x, (2, 1, 0), lo, hi)
xf = x.to(torch.float32)
for ax in (2, 1, 0):
    IO._rg_pass(xf, ax, 2.0, 0)
torch.cuda.synchronize()
PY
summ ncu_filters_256
ls -la $R
# keep the reports that fit (64 MiB cap on gpurun_out): smallest first
used=$(du -sm $O | cut -f1)
for f in $(ls -S -r $R/*.ncu-rep); do sz=$(( $(stat -c %s $f) / 1048576 + 1 )); if [ $((used + sz)) -lt 55 ]; then cp $f $O/; used=$((used + sz)); fi; done
ls -la $O
