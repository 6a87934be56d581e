# A/B of library variants on one box: scripts/ab_variants.sh SIZE KIND [variant names...]
size=$1; kind=$2; shift 2
echo "== default"; python scripts/quick_time.py $size $kind 2>&1 | grep -E "glcm|suite"
for v in "$@"; do echo "== $v"; B200_RADIOMICS_LIB=pyradiomics_b200/variants/lib$v.so python scripts/quick_time.py $size $kind 2>&1 | grep -E "glcm|suite"; done
