set -x
O=gpurun_out/r02_call14; mkdir -p $O
V=pyradiomics_b200/variants
{
for sp in 0 1 2 3; do echo "default lib, B200_GLCM_SPLIT=$sp"; for k in uniform smooth; do B200_GLCM_SPLIT=$sp python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done
for sp in 1 3; do echo "segsum1 lib, B200_GLCM_SPLIT=$sp"; for k in uniform smooth; do B200_RADIOMICS_LIB=$V/libsegsum1.so B200_GLCM_SPLIT=$sp python scripts/quick_time.py 256 $k | grep -E "glcm"; done; done
} 2>&1 | grep -v "^+" | tee $O/quick_time_split.txt
B200_GLCM_SPLIT=3 python scripts/diag_mcc.py 96 2>&1 | cut -c1-160 | tee $O/diag_split3.txt
B200_GLCM_SPLIT=3 bash scripts/launch_list.sh 256 smooth /tmp/l3.csv 50 > $O/launches_smooth_split3.txt 2>&1
B200_GLCM_SPLIT=0 bash scripts/launch_list.sh 256 smooth /tmp/l0.csv 25 > $O/launches_smooth_split0.txt 2>&1
ls -la $O
