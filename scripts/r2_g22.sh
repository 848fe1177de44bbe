set -x
V=$PWD/racon_gpu_b200/variants
L=gpurun_out/r2_g22_ab.log
for v in base_old cur new base_old new; do
  if [ $v = new ]; then unset B200POA_LIB; else export B200POA_LIB=$V/libb200poa_$v.so; fi
  echo "== $v banded" >> $L
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
done
for v in base_old new; do
  if [ $v = new ]; then unset B200POA_LIB; else export B200POA_LIB=$V/libb200poa_$v.so; fi
  echo "== $v full" >> $L
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 >> $L 2>&1
done
unset B200POA_LIB
echo "== new B_banded" >> $L
timeout 300 python scripts/profile_run.py --windows 4096 --length 1024 --depth 64 --err 0.12 --max-seq 1279 --banded 1 --launches 2 --mem-gb 100 >> $L 2>&1
