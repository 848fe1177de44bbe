set -x
L=gpurun_out/r2_g21.log
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_g21_tests.log
for band in 0 1 adaptive; do
  B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fastq_500 --band $band >> $L 2>&1
done
B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fasta_500 --band 1 >> $L 2>&1
B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fastq_1000 --band 1 --copies 160 >> $L 2>&1
timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
