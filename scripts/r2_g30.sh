set -x
timeout 600 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_g30_aln_tests.log
L=gpurun_out/r2_g30_aln_bench.log; : > $L
for rep in 1 8 16 32 64; do echo "== rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g30_launches_rep64.csv python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 > /dev/null 2>&1
