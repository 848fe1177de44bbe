set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_device or arena" 2>&1 | tail -5 > gpurun_out/r2_g15_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_g15_bench_n2.json 2> gpurun_out/r2_g15_bench_n2.err
timeout 900 python scripts/multi_device_bench.py --batches 2 > gpurun_out/r2_g15_multidev.log 2>&1
