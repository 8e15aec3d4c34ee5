# One measurement pass on the GPU box (bash tools/measure_round.sh [tag], default tag r02): parity suite, smoke, ncu evidence
# (launch lists with DRAM bytes + --set full captures of the dominant kernel of c2 / c3 / c5, each only after the plain run of
# the same command has exited 0), profiles/roofline_latest.json from them, then both bench arms (the bench line reads the
# profile-time traffic / issue figures from that file), every BASELINE config briefly, the output stage. Results land in
# gpurun_out/; the summaries worth keeping are copied to profiles/ by hand afterwards.
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/pytest_$TAG.log 2>&1; tail -14 gpurun_out/pytest_$TAG.log
grep -hE "IMAGE_STATS|C3_CRN|CRN_FULL|FAILED|^E  " gpurun_out/pytest_$TAG.log | cut -c1-330 > gpurun_out/pytest_${TAG}_stats.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for w in c2:2 c3:16 c5:1; do
  n=${w%%:*}; p=${w##*:}
  timeout 120 python tools/profile_run.py $n $p > gpurun_out/plain_$n.log 2>&1 && \
  timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_$n.csv python tools/profile_run.py $n $p > gpurun_out/ncu_l_$n.log 2>&1
  cat gpurun_out/plain_$n.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/prof_${TAG}_c2 python tools/profile_run.py c2 2 > gpurun_out/ncu_f_c2.log 2>&1; tail -1 gpurun_out/ncu_f_c2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_mesh -s 8 -c 2 -o gpurun_out/prof_${TAG}_c3 python tools/profile_run.py c3 16 > gpurun_out/ncu_f_c3.log 2>&1; tail -1 gpurun_out/ncu_f_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/prof_${TAG}_c5 python tools/profile_run.py c5 1 > gpurun_out/ncu_f_c5.log 2>&1; tail -1 gpurun_out/ncu_f_c5.log
for n in c2 c3 c5; do
  ncu -i gpurun_out/prof_${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$n.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_$n.csv 2>/dev/null
done
python tools/ncu_source_lines.py gpurun_out/src_c2.csv "k_shade<(int)1" 60 > gpurun_out/ncu_${TAG}_k_shade_next_source_top60.txt
python tools/ncu_source_lines.py gpurun_out/src_c2.csv "k_shade<(int)2" 60 > gpurun_out/ncu_${TAG}_k_shade_fused_source_top60.txt
python tools/ncu_source_lines.py gpurun_out/src_c3.csv "k_extend_mesh" 60 > gpurun_out/ncu_${TAG}_k_extend_mesh_source_top60.txt
python tools/ncu_source_lines.py gpurun_out/src_c5.csv "k_shade<(int)2" 40 > gpurun_out/ncu_${TAG}_k_shade_light_bvh_fused_source_top40.txt
python tools/ncu_source_lines.py gpurun_out/src_c5.csv "k_shade<(int)1" 40 > gpurun_out/ncu_${TAG}_k_shade_light_bvh_next_source_top40.txt
rm -f gpurun_out/src_c2.csv gpurun_out/src_c3.csv gpurun_out/src_c5.csv gpurun_out/prof_${TAG}_c*.ncu-rep
python tools/make_roofline.py c2=gpurun_out/launches_${TAG}_c2.csv,gpurun_out/ncu_${TAG}_c2_raw.csv c3=gpurun_out/launches_${TAG}_c3.csv,gpurun_out/ncu_${TAG}_c3_raw.csv c5=gpurun_out/launches_${TAG}_c5.csv,gpurun_out/ncu_${TAG}_c5_raw.csv > gpurun_out/roofline_summary.txt 2>&1; tail -3 gpurun_out/roofline_summary.txt
cp profiles/roofline_latest.json gpurun_out/roofline_latest.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_${TAG}_reference.json
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2>> gpurun_out/bench_err.log; tail -c 400 gpurun_out/bench_${TAG}.json
timeout 600 python tools/run_configs.py c1,c2,c3,c3_tree,c4,c5_100,c5 > gpurun_out/configs_$TAG.jsonl 2>&1; cut -c1-200 gpurun_out/configs_$TAG.jsonl
for w in c3 c4 c5 c1; do timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-c4 >> gpurun_out/bench_${TAG}_other_workloads.jsonl 2>> gpurun_out/bench_err.log; done
timeout 300 python tools/bench_output.py > gpurun_out/output_stage_$TAG.jsonl 2>&1; timeout 300 python tools/bench_output.py --sources 20000 --no-cpu >> gpurun_out/output_stage_$TAG.jsonl 2>&1; cut -c1-200 gpurun_out/output_stage_$TAG.jsonl
tail -5 gpurun_out/bench_err.log
