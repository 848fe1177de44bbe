set -x
L=gpurun_out/r2_g49_poa_occupancy.log; : > $L
for v in - poa16 poa20 poa22 -; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  echo "== $v banded" >> $L
  timeout 200 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 2>&1 | grep -E "launch ms|info" | cut -c1-260 >> $L
  echo "== $v full" >> $L
  timeout 200 python scripts/profile_run.py --windows 10000 --banded 0 --launches 2 --mem-gb 64 2>&1 | grep -E "launch ms" >> $L
done
