#!/bin/bash
# final ncu evidence of the round: launch list of a cfg3 step + full captures of the kernels this round rebuilt
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_step.csv python tools/profile_step.py --steps 2 > gpurun_out/r02n_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paf_ -c 2 -f -o gpurun_out/r02_final_paf python tools/profile_step.py --steps 1 > gpurun_out/r02n_ncu_paf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 2 -f -o gpurun_out/r02_final_halo python tools/profile_step.py --steps 1 > gpurun_out/r02n_ncu_halo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stem7 -c 1 -f -o gpurun_out/r02_final_stem7 python tools/profile_cfg.py --graph resnet50_lw_openpose --batch 32 --steps 1 > gpurun_out/r02n_ncu_stem7.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_cfg2.csv python tools/profile_cfg.py --steps 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
