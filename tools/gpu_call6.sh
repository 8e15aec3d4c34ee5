# Round-2 GPU call 6: whole parity suite, then ncu evidence for every kernel family: launch lists (time + DRAM bytes) of
# c2 / c3 / c5, --set full captures of the fused shade kernels (c2), the mesh traversal kernel (c3) and the light-LBVH
# shade kernels (c5), each after the plain run of the same command has exited 0.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --durations=10 > gpurun_out/pytest6.log 2>&1; tail -16 gpurun_out/pytest6.log
grep -hE "IMAGE_STATS|C3_CRN|CRN_FULL|FAILED|^E  " gpurun_out/pytest6.log | cut -c1-330 | head -40
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for w in c2:2 c3:1 c5:1; do
  n=${w%%:*}; p=${w##*:}
  timeout 120 python tools/profile_run.py $n $p > gpurun_out/plain_$n.log 2>&1 && \
  timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_$n.csv python tools/profile_run.py $n $p > gpurun_out/ncu_l_$n.log 2>&1
  cat gpurun_out/plain_$n.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/prof_r02_c2 python tools/profile_run.py c2 2 > gpurun_out/ncu_f_c2.log 2>&1; tail -1 gpurun_out/ncu_f_c2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_mesh -s 8 -c 2 -o gpurun_out/prof_r02_c3 python tools/profile_run.py c3 1 > gpurun_out/ncu_f_c3.log 2>&1; tail -1 gpurun_out/ncu_f_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/prof_r02_c5 python tools/profile_run.py c5 1 > gpurun_out/ncu_f_c5.log 2>&1; tail -1 gpurun_out/ncu_f_c5.log
for n in c2 c3 c5; do ncu -i gpurun_out/prof_r02_$n.ncu-rep --page raw --csv > gpurun_out/ncu_r02_${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
