set -x
O=gpurun_out/r02_call11; mkdir -p $O
V=pyradiomics_b200/variants
for v in "" lzacc1 eigstatic tile4k; do echo "variant=$v"; if [ -n "$v" ]; then export B200_RADIOMICS_LIB=$V/lib$v.so; else unset B200_RADIOMICS_LIB; fi; for k in uniform smooth; do python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done 2>&1 | grep -v "^+" | tee $O/quick_time_variants.txt
B200_RADIOMICS_LIB=$V/libeigstatic.so python scripts/diag_mcc.py 96 2>&1 | cut -c1-160 | tee $O/diag_eigstatic.txt
unset B200_RADIOMICS_LIB
timeout 600 ncu --set full --clock-control none -k regex:glcm_fast -s 25 -c 5 -o $O/ncu_glcm_256_uniform python scripts/prof_glcm.py 256 uniform glcm > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:glcm_fast -s 25 -c 5 -o $O/ncu_glcm_256_smooth python scripts/prof_glcm.py 256 smooth glcm > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"seg_tile|seg_glrlm_ends|ccl_merge" -c 5 -o $O/ncu_segment_256 python scripts/prof_segment.py 256 smooth 1 > /dev/null 2>&1
python - <<'PY' > $O/filters_time.txt 2>&1
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from pyradiomics_b200 import imageoperations as IO
x = torch.randn((256, 256, 256), device="cuda", dtype=torch.float64)
lo, hi = IO.wavelet_filters("coif1")
def t(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("swt3d fused 256^3 ms", t(lambda: IO.swt_level1_device(x, (2, 1, 0), lo, hi)))
xf = x.to(torch.float32)
print("LoG sigma 2 256^3 ms", t(lambda: IO.log_filter_device(xf, 2.0, (1.0, 1.0, 1.0))))
for ax in (0, 1, 2):
    print("recursive gauss axis", ax, "ms", t(lambda: IO._rg_pass(xf, ax, 2.0, 0)))
print("bin 256^3 ms", t(lambda: IO.bin_image_device(x, None, binWidth=0.25)))
PY
cat $O/filters_time.txt
timeout 300 ncu --set full --clock-control none -k regex:"swt3d|recursive_gauss" -c 4 -o $O/ncu_filters_256 python - <<'PY' > /dev/null 2>&1
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
ls -la $O
for f in $O/*.ncu-rep; do ncu -i $f --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__grid_size,launch__block_size > ${f%.ncu-rep}.csv 2>&1; done
ls -la $O
