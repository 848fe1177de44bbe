set -x
timeout 900 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_g25_aln_tests.log
timeout 300 python scripts/aln_bench.py --rep 8 --iters 2 > gpurun_out/r2_g25_aln_bench.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_aligner.py -x -q -m gpu -k "argument or live" 2>&1 | tail -15 > gpurun_out/r2_g25_sanitizer.log
