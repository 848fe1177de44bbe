set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_g3_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_g3_bench.json 2> gpurun_out/r2_g3_bench.err
python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 > gpurun_out/r2_g3_afull.log 2>&1
python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 > gpurun_out/r2_g3_abanded.log 2>&1
