set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_g45_smoke.log 2>&1
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_g45_tests.log
timeout 900 python bench.py > gpurun_out/r2_g45_bench.json 2> gpurun_out/r2_g45_bench.err
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_g45_sanitizer_full.log python -m pytest tests/test_gpu_aligner.py -x -q -m gpu -k "not saturated and not pool and not guesses" > gpurun_out/r2_g45_sanitizer_pytest.log 2>&1
grep -n "=========" gpurun_out/r2_g45_sanitizer_full.log | grep -v "Host Frame\|Saved host\|=========$" | head -20 > gpurun_out/r2_g45_sanitizer_summary.log
