set -x
L=gpurun_out/r2_g24.log
timeout 900 python -m pytest tests/test_gpu_msa.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_g24_msa_tests.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_msa.py 2>&1 | tail -8 > gpurun_out/r2_g24_tests.log
for i in 1 2; do
  timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
done
timeout 300 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 >> $L 2>&1
