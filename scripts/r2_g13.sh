set -x
python scripts/e2e_probe.py > gpurun_out/r2_g13_e2e.log 2>&1
B200POA_E2E_TIMERS=1 python - > gpurun_out/r2_g13_e2e_timers.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from racon_gpu_b200 import api
from racon_gpu_b200.windows import synth_windows
b = synth_windows(10000, 500, 32, 0.15, seed=12345)
pol = api.Polisher(devices=[0], batches_per_device=4, mem_per_batch=12 << 30, banded=True)
for _ in range(3): pol.polish(b, max_windows_per_round=2500)
PY
for v in "" tma; do
  if [ -n "$v" ]; then export B200POA_LIB=$PWD/racon_gpu_b200/variants/libb200poa_$v.so; fi
  python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 > gpurun_out/r2_g13_abanded_${v:-base}.log 2>&1
done
