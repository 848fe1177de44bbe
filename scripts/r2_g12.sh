set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_g12_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_g12_bench.json 2> gpurun_out/r2_g12_bench.err
timeout 900 python bench.py --workload A_banded_1M --batches 8 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_g12_bench_1M.json 2> gpurun_out/r2_g12_bench_1M.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_g12_bench_ref.json 2> gpurun_out/r2_g12_bench_ref.err
