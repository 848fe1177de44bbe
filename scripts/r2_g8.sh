python scripts/dbg_adaptive3.py > gpurun_out/r2_dbg3_plain.log 2>&1
timeout 500 compute-sanitizer --tool initcheck --print-limit 8 python scripts/dbg_adaptive3.py > gpurun_out/r2_dbg3_initcheck.log 2>&1
timeout 500 compute-sanitizer --tool racecheck --print-limit 8 python scripts/dbg_adaptive3.py > gpurun_out/r2_dbg3_racecheck.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 8 python scripts/dbg_adaptive3.py > gpurun_out/r2_dbg3_memcheck.log 2>&1
