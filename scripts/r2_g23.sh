set -x
L=gpurun_out/r2_g23.log
timeout 900 python -m pytest tests/test_gpu_real_data.py tests/test_gpu_parity.py -x -q -m gpu -k "real or lambda or partial or span or cudapoa" 2>&1 | tail -5 > gpurun_out/r2_g23_tests.log
for band in 0 1 adaptive; do
  B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fastq_500 --band $band >> $L 2>&1
done
B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fastq_1000 --band 1 --copies 160 >> $L 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:poa_window_kernel -c 1 -o gpurun_out/r2_abanded_v43 python scripts/profile_run.py --windows 10000 --banded 1 --launches 1 --mem-gb 64 > gpurun_out/r2_g23_ncu.log 2>&1
