set -x
mkdir -p gpurun_out
timeout 120 python tools/profile_run.py c3 1 > gpurun_out/plain_c3.log 2>&1; cat gpurun_out/plain_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_mesh -s 8 -c 2 -o gpurun_out/prof_r02b_c3 python tools/profile_run.py c3 1 > gpurun_out/ncu_f_c3.log 2>&1; tail -1 gpurun_out/ncu_f_c3.log
timeout 120 python tools/profile_run.py c5 1 > gpurun_out/plain_c5.log 2>&1; cat gpurun_out/plain_c5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/prof_r02b_c5 python tools/profile_run.py c5 1 > gpurun_out/ncu_f_c5.log 2>&1; tail -1 gpurun_out/ncu_f_c5.log
for n in c3 c5; do ncu -i gpurun_out/prof_r02b_$n.ncu-rep --page raw --csv > gpurun_out/ncu_r02b_${n}_raw.csv 2>/dev/null; done
