set -x
L=gpurun_out/r2_g43_aln_long.log; : > $L
echo "== synthetic 1500 x 30 kb, 12 % (view)" >> $L
timeout 600 python scripts/aln_bench.py --synthetic 1500,30000,0.12 --iters 2 --cpu-sample 1 --view 1 --mem-gb 64 >> $L 2>&1
echo "== synthetic 400 x 60 kb, 10 % (view)" >> $L
timeout 600 python scripts/aln_bench.py --synthetic 400,60000,0.10 --iters 2 --cpu-sample 1 --view 1 --mem-gb 64 >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_leaf_kernel -c 1 -f -o gpurun_out/r2_g43_aln_leaf_rep64 python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > gpurun_out/r2_g43_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_split_kernel -c 2 -f -o gpurun_out/r2_g43_aln_split_rep64 python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > gpurun_out/r2_g43_ncu2.log 2>&1
