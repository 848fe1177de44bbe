set -x
V=$PWD/racon_gpu_b200/variants
L=gpurun_out/r2_g18_ab.log
for v in base_old dirmap1 dirmap2; do
  export B200POA_LIB=$V/libb200poa_$v.so
  echo "== $v banded" >> $L
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
done
export B200POA_LIB=$V/libb200poa_dirmap2sub.so
echo "== dirmap2sub banded" >> $L
B200POA_PHASE_TIMERS=1 timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
for v in base_old dirmap2; do
  export B200POA_LIB=$V/libb200poa_$v.so
  echo "== $v full" >> $L
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 >> $L 2>&1
done
unset B200POA_LIB
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_g18_tests.log
