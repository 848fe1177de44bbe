set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_g4_tests.log
python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 > gpurun_out/r2_g4_abanded.log 2>&1
python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 > gpurun_out/r2_g4_afull.log 2>&1
B200POA_PHASE_TIMERS=1 python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 > gpurun_out/r2_g4_abanded_phases.log 2>&1
python scripts/profile_run.py --windows 7104 --banded 1 --length 1024 --depth 64 --err 0.12 --max-seq 1279 --launches 2 --mem-gb 64 > gpurun_out/r2_g4_b.log 2>&1
