set -x
L=gpurun_out/r2_g20.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:poa_window_kernel -c 1 -o gpurun_out/r2_abanded_v42 python scripts/profile_run.py --windows 10000 --banded 1 --launches 1 --mem-gb 64 > gpurun_out/r2_g20_ncu.log 2>&1
for band in 0 1 adaptive; do
  B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fastq_500 --band $band >> $L 2>&1
done
B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fasta_500 --band 1 >> $L 2>&1
B200POA_PHASE_TIMERS=1 timeout 600 python scripts/real_data_bench.py --case fastq_1000 --band 1 --copies 160 >> $L 2>&1
timeout 300 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 >> $L 2>&1
