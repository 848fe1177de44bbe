set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g35_aln_tests.log
L=gpurun_out/r2_g35_aln_bench.log; : > $L
for v in - unroll16_base - unroll16_base; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  for rep in 8 64; do echo "== $v rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1; done
done
