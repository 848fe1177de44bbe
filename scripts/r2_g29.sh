set -x
timeout 300 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_g29_aln_tests.log
L=gpurun_out/r2_g29_aln_bench.log; : > $L
for rep in 1 8 64; do echo "== rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_split_kernel -c 1 -f -o gpurun_out/r2_g29_aln_split_rep64 python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > gpurun_out/r2_g29_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_leaf_kernel -c 1 -f -o gpurun_out/r2_g29_aln_leaf_rep64 python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > gpurun_out/r2_g29_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aln_split_team_kernel -c 1 -f -o gpurun_out/r2_g29_aln_team_rep8 python scripts/aln_bench.py --rep 8 --iters 1 --cpu-sample 0 > gpurun_out/r2_g29_ncu3.log 2>&1
