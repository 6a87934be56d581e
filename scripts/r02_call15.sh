set -x
O=gpurun_out/r02_call15; mkdir -p $O
python scripts/prof_host.py 256 256 > $O/prof_host.txt 2>&1
tail -5 $O/prof_host.txt
# one launch each of GLRLM / NGTDM / phase A with source lines
for c in glrlm ngtdm; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:"glrlm_fast_kernel|small_fast_kernel" -s 1 -c 1 -o /tmp/ncu_$c python scripts/prof_glcm.py 256 uniform $c > $O/ncu_$c.log 2>&1
  python scripts/summarize_ncu.py /tmp/ncu_$c.ncu-rep > $O/ncu_${c}_256_uniform.txt 2>&1
  ncu -i /tmp/ncu_$c.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$c.csv 2>$O/src_$c.err
  python scripts/src_lines.py /tmp/src_$c.csv 60 >> $O/ncu_${c}_256_uniform.txt 2>&1
done
ls -la $O /tmp/*.ncu-rep
