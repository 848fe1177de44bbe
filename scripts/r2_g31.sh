set -x
timeout 600 python -m pytest tests/test_gpu_aligner.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_g31_aln_tests.log
L=gpurun_out/r2_g31_aln_bench.log; : > $L
for rep in 1 8 64; do echo "== rep $rep" >> $L; timeout 300 python scripts/aln_bench.py --rep $rep --iters 3 --cpu-sample 0 >> $L 2>&1; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_g31_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_g31_bench.json 2> gpurun_out/r2_g31_bench.err
timeout 1800 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_aligner.py 2>&1 | tail -8 > gpurun_out/r2_g31_tests.log
