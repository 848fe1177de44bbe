set -x
timeout 600 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g26_aln_tests.log
L=gpurun_out/r2_g26_aln_bench.log; : > $L
for v in - aln12 aln20 aln24 - aln24; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  for rep in 8 64; do
    echo "== $v rep $rep" >> $L
    timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1
  done
done
unset B200POA_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_split_kernel -c 1 -f -o gpurun_out/r2_g26_aln_split python scripts/aln_bench.py --rep 32 --iters 1 --cpu-sample 0 > gpurun_out/r2_g26_ncu_split.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_leaf_kernel -c 1 -f -o gpurun_out/r2_g26_aln_leaf python scripts/aln_bench.py --rep 32 --iters 1 --cpu-sample 0 > gpurun_out/r2_g26_ncu_leaf.log 2>&1
