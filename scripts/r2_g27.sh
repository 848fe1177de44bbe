set -x
timeout 300 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_g27_aln_tests.log
L=gpurun_out/r2_g27_aln_bench.log; : > $L
for rep in 1 8 64; do
  echo "== rep $rep" >> $L
  timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1
done
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_aligner.py -x -q -m gpu -k "real or live" 2>&1 | tail -6 > gpurun_out/r2_g27_sanitizer.log
