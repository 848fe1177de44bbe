set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_g41_smoke.log 2>&1
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_g41_tests.log
timeout 900 python bench.py > gpurun_out/r2_g41_bench.json 2> gpurun_out/r2_g41_bench.err
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_aligner.py -x -q -m gpu -k "not saturated and not pool" 2>&1 | tail -6 > gpurun_out/r2_g41_sanitizer.log
