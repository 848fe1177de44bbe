timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_g42_sanitizer_full.log python -m pytest tests/test_gpu_aligner.py -x -q -m gpu -k "not saturated and not pool" > gpurun_out/r2_g42_pytest.log 2>&1
grep -n "=========" gpurun_out/r2_g42_sanitizer_full.log | grep -v "Host Frame\|Saved host\|=========$" | head -40 > gpurun_out/r2_g42_sanitizer_summary.log
