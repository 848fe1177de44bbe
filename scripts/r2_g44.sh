set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_g44_aln_tests.log
L=gpurun_out/r2_g44_aln_bench.log; : > $L
for g in 0 -1; do
  for rep in 8 64; do echo "== guess $g rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 --view 1 --guess $g >> $L 2>&1; done
  echo "== guess $g synthetic 1500 x 30 kb, 12 %" >> $L
  timeout 600 python scripts/aln_bench.py --synthetic 1500,30000,0.12 --iters 2 --cpu-sample 0 --view 1 --mem-gb 64 --guess $g >> $L 2>&1
  echo "== guess $g synthetic 400 x 60 kb, 10 %" >> $L
  timeout 600 python scripts/aln_bench.py --synthetic 400,60000,0.10 --iters 2 --cpu-sample 0 --view 1 --mem-gb 64 --guess $g >> $L 2>&1
done
