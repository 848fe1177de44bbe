set -x
V=$PWD/racon_gpu_b200/variants
L=gpurun_out/r2_g19_ab.log
for v in base_old ef du6 du8 pf1 pf2 pf3 base_old; do
  export B200POA_LIB=$V/libb200poa_$v.so
  echo "== $v banded" >> $L
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
done
for v in base_old pf1 pf2; do
  export B200POA_LIB=$V/libb200poa_$v.so
  echo "== $v full" >> $L
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 >> $L 2>&1
done
