set -x
python -m pytest tests/test_gpu_real_data.py -x -q 2>&1 | tail -15 > gpurun_out/r2_realdata.log
B200POA_PHASE_TIMERS=1 python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 > gpurun_out/r2_prof_afull.log 2>&1
B200POA_PHASE_TIMERS=1 python scripts/profile_run.py --windows 2000 --banded 1 --length 1024 --depth 64 --err 0.12 --max-seq 1279 --launches 3 --mem-gb 60 > gpurun_out/r2_prof_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_window -s 1 -c 1 -o gpurun_out/r2_afull_v35 -f python scripts/profile_run.py --windows 3552 --banded 0 --launches 2 > gpurun_out/r2_ncu_afull.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_window -s 1 -c 1 -o gpurun_out/r2_b_v35 -f python scripts/profile_run.py --windows 1000 --banded 1 --length 1024 --depth 64 --err 0.12 --max-seq 1279 --launches 2 --mem-gb 60 > gpurun_out/r2_ncu_b.log 2>&1
