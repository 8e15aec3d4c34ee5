# Round-2 GPU call 2: parity suite, the bench line (with the C4 sub-record), ncu launch list + full capture of the fused
# shade kernels, many-light A/B.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/pytest2.log 2>&1; tail -3 gpurun_out/pytest2.log
grep -E "IMAGE_STATS|FAILED|^E  " gpurun_out/pytest2.log | cut -c1-420 | head -60
python bench.py --steps 24 --warmup 3 --with-c4 --c4-steps 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 1500 gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
python tools/ab_r02.py configs base,wide c3,c3_tree > gpurun_out/ab2_mesh.log 2>&1; cat gpurun_out/ab2_mesh.log
python tools/ab_r02.py configs base,oneq c5_100,c5 > gpurun_out/ab2_lights.log 2>&1; cat gpurun_out/ab2_lights.log
python tools/ab_r02.py run base,r10,nearr1 c2 > gpurun_out/ab2_c2.log 2>&1; cat gpurun_out/ab2_c2.log
python tools/run_configs.py c1,c2,c3,c3_tree,c5_100,c5 > gpurun_out/configs2.jsonl 2>&1
# ncu: launch list with DRAM bytes, then the two dominant launches in full
python tools/profile_run.py cornell 1024 1024 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python tools/profile_run.py cornell 1024 1024 2 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -o gpurun_out/prof_shade python tools/profile_run.py cornell 1024 1024 2 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
ncu -i gpurun_out/prof_shade.ncu-rep --page raw --csv > gpurun_out/ncu_r02_k_shade_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_shade.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_shade.csv 2>/dev/null
python tools/ncu_source_lines.py gpurun_out/src_shade.csv "k_shade<1" 60 > gpurun_out/ncu_r02_k_shade_next_source_top60.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/src_shade.csv "k_shade<2" 60 > gpurun_out/ncu_r02_k_shade_fused_source_top60.txt 2>&1
head -30 gpurun_out/ncu_r02_k_shade_fused_source_top60.txt
ls -la gpurun_out/; rm -f gpurun_out/src_shade.csv
du -sh gpurun_out
