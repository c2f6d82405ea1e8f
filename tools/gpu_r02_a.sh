#!/bin/bash
# round-2 evidence run (one gpurun call): parser phase stamps, multicast A/B, ncu launch list + full captures
mkdir -p gpurun_out
python tools/paf_phases.py > gpurun_out/r02_paf_phases.txt 2>&1
HPB_SWAP_MC=1 timeout 300 python bench.py --no-extra --no-tf32-line --no-cpu-baseline --steps 30 > gpurun_out/r02_bench_swapmc.json 2> gpurun_out/r02_bench_swapmc.err
cp gpurun_out/bench_layers_cfg3_f16_n1.json gpurun_out/r02_layers_swapmc.json
timeout 300 python bench.py --no-extra --no-tf32-line --no-cpu-baseline --steps 30 > gpurun_out/r02_bench_nomc.json 2> gpurun_out/r02_bench_nomc.err
cp gpurun_out/bench_layers_cfg3_f16_n1.json gpurun_out/r02_layers_nomc.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv python tools/profile_step.py --steps 2 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paf_ -c 2 -f -o gpurun_out/r02_paf python tools/profile_step.py --steps 1 > gpurun_out/r02_ncu_paf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_swap -s 6 -c 1 -f -o gpurun_out/r02_swap python tools/profile_step.py --steps 1 > gpurun_out/r02_ncu_swap.log 2>&1
ls -la gpurun_out | tail -20
