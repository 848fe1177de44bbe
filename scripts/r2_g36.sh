set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g36_aln_tests.log
L=gpurun_out/r2_g36_aln_bench.log; : > $L
for v in - walk1 - walk1; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  for rep in 8 64; do echo "== $v rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1; done
done
unset B200POA_LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g36_launches_rep64.csv python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > /dev/null 2>&1
