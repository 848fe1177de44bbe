set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g46_launches_60kb.csv python scripts/aln_bench.py --synthetic 400,60000,0.10 --iters 1 --cpu-sample 0 --view 1 --mem-gb 64 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g46_launches_rep64.csv python scripts/aln_bench.py --rep 64 --iters 1 --cpu-sample 0 --view 1 > /dev/null 2>&1
