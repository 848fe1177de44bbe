timeout 600 python scripts/aln_fuzz_gpu.py 150 > gpurun_out/r2_g48_fuzz.log 2>&1
