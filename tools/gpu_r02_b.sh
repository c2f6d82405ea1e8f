#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_paf_gpu.py tests/test_handoff.py tests/test_pipeline_pool.py -x -q > gpurun_out/r02b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_tests.log)
tail -15 gpurun_out/r02b_tests.log
python tools/paf_phases.py > gpurun_out/r02b_paf_phases.txt 2>&1; cat gpurun_out/r02b_paf_phases.txt
python tools/time_paf.py > gpurun_out/r02b_time_paf.txt 2>&1; cat gpurun_out/r02b_time_paf.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:paf_ --log-file gpurun_out/r02b_paf_launches.csv python tools/profile_step.py --steps 2 > /dev/null 2>&1
grep paf_ gpurun_out/r02b_paf_launches.csv | cut -d, -f5,15- | tail -4
