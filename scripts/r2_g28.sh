set -x
L=gpurun_out/r2_g28_aln_bench.log; : > $L
for v in - team8 team2 waves16; do
  if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB=racon_gpu_b200/variants/libb200poa_$v.so; fi
  for rep in 1 8 64; do
    echo "== $v rep $rep" >> $L
    timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1
  done
done
unset B200POA_LIB
for rep in 8 64; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g28_launches_rep$rep.csv python scripts/aln_bench.py --rep $rep --iters 1 --cpu-sample 0 > /dev/null 2>&1
done
