set -x
L=gpurun_out/r2_g39_aln_bench.log; : > $L
for v in - mb12 mb20 mb24 -; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  for rep in 8 64; do echo "== $v rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 --view 1 >> $L 2>&1; done
done
