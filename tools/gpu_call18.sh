# c3 profile at the bench's batch size (16 passes = one 2^25-path batch), then roofline_latest.json again
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 python tools/profile_run.py c3 16 > gpurun_out/plain_c3.log 2>&1 && timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_c3.csv python tools/profile_run.py c3 16 > gpurun_out/ncu_l_c3.log 2>&1
cat gpurun_out/plain_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_mesh -s 8 -c 2 -o gpurun_out/prof_r02_c3 python tools/profile_run.py c3 16 > gpurun_out/ncu_f_c3.log 2>&1; tail -1 gpurun_out/ncu_f_c3.log
ncu -i gpurun_out/prof_r02_c3.ncu-rep --page raw --csv > gpurun_out/ncu_r02_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02_c3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_c3.csv 2>/dev/null
python tools/ncu_source_lines.py gpurun_out/src_c3.csv "k_extend_mesh" 60 > gpurun_out/ncu_r02_k_extend_mesh_source_top60.txt
rm -f gpurun_out/src_c3.csv gpurun_out/prof_r02_c3.ncu-rep
python tools/make_roofline.py c2=profiles/launches_r02_c2.csv,profiles/ncu_r02_k_shade_raw.csv c3=gpurun_out/launches_r02_c3.csv,gpurun_out/ncu_r02_c3_raw.csv c5=profiles/launches_r02_c5.csv,profiles/ncu_r02_k_shade_light_bvh_raw.csv | tail -25
cp profiles/roofline_latest.json gpurun_out/roofline_latest.json
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-c4 > gpurun_out/bench_r02_c3.json 2>> gpurun_out/bench_err.log; tail -c 1200 gpurun_out/bench_r02_c3.json
