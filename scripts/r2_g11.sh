timeout 400 python scripts/dbg_adaptive6.py > gpurun_out/r2_dbg6_fence.log 2>&1
B200POA_LIB=$PWD/racon_gpu_b200/variants/libb200poa_nofence.so timeout 400 python scripts/dbg_adaptive6.py > gpurun_out/r2_dbg6_nofence.log 2>&1
