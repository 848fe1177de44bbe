set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_g5_tests.log
for rep in 1 2; do
python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 > gpurun_out/r2_g5_abanded_$rep.log 2>&1
B200POA_LIB=$PWD/racon_gpu_b200/variants/libb200poa_mb28.so python scripts/profile_run.py --windows 10000 --banded 1 --launches 3 --mem-gb 64 > gpurun_out/r2_g5_abanded_mb28_$rep.log 2>&1
done
python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 > gpurun_out/r2_g5_afull.log 2>&1
B200POA_LIB=$PWD/racon_gpu_b200/variants/libb200poa_mb28.so python scripts/profile_run.py --windows 10000 --banded 0 --launches 3 --mem-gb 64 > gpurun_out/r2_g5_afull_mb28.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_window -s 1 -c 1 -o gpurun_out/r2_abanded_v41 -f python scripts/profile_run.py --windows 10000 --banded 1 --launches 2 --mem-gb 64 > gpurun_out/r2_g5_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_window -s 1 -c 1 -o gpurun_out/r2_afull_v41 -f python scripts/profile_run.py --windows 10000 --banded 0 --launches 2 --mem-gb 64 > gpurun_out/r2_g5_ncu2.log 2>&1
cp racon_gpu_b200/libb200poa.so gpurun_out/libb200poa_v41.so
