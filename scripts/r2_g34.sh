set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g34_aln_tests.log
L=gpurun_out/r2_g34_aln_bench.log; : > $L
for v in - unroll1 unroll4 - unroll1; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  for rep in 8 64; do echo "== $v rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1; done
done
unset B200POA_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_split_kernel -c 1 -f -o gpurun_out/r2_g34_aln_split_rep64 python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > gpurun_out/r2_g34_ncu1.log 2>&1
