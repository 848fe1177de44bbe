set -x
V=$PWD/racon_gpu_b200/variants
for rep in 1 2; do
for v in base tb2x6 tb3x4 tb2x3 tb1x4 u8; do
  if [ $v = base ]; then unset B200POA_LIB; else export B200POA_LIB=$V/libb200poa_$v.so; fi
  echo "== $v banded rep $rep" >> gpurun_out/r2_g17_ab.log
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> gpurun_out/r2_g17_ab.log 2>&1
done
done
for v in base tb2x6 tb3x4 u8; do
  if [ $v = base ]; then unset B200POA_LIB; else export B200POA_LIB=$V/libb200poa_$v.so; fi
  echo "== $v full" >> gpurun_out/r2_g17_ab.log
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 >> gpurun_out/r2_g17_ab.log 2>&1
done
unset B200POA_LIB
echo "== base banded phases" >> gpurun_out/r2_g17_ab.log
B200POA_PHASE_TIMERS=1 timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> gpurun_out/r2_g17_ab.log 2>&1
echo "== base full phases" >> gpurun_out/r2_g17_ab.log
B200POA_PHASE_TIMERS=1 timeout 300 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 >> gpurun_out/r2_g17_ab.log 2>&1
export B200POA_LIB=$PWD/racon_gpu_b200/variants/libb200poa_sub.so
echo "== sub banded subtimers" >> gpurun_out/r2_g17_ab.log
B200POA_PHASE_TIMERS=1 timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> gpurun_out/r2_g17_ab.log 2>&1
