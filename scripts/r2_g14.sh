set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_g14_tests.log
for bt in 1 2 4; do timeout 600 python bench.py --steps 5 --warmup 3 --batches $bt --no-cpu-baseline --no-extra > gpurun_out/r2_g14_bench_b$bt.json 2> gpurun_out/r2_g14_bench_b$bt.err; done
timeout 900 python bench.py --workload A_banded_1M --batches 8 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_g14_bench_1M.json 2> gpurun_out/r2_g14_bench_1M.err
