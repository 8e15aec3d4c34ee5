# One measurement pass on the GPU box (bash tools/measure_round.sh): tests, smoke, both bench arms, every BASELINE config,
# the ncu launch list (+DRAM bytes) of one Cornell pass, the output-stage timing. Results land in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_r01.json
python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_r01_reference.json
python tools/run_configs.py c1,c2,c3,c3_tree,c5_100,c5 > gpurun_out/configs.jsonl 2>&1; cut -c1-260 gpurun_out/configs.jsonl
python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2>> gpurun_out/bench_err.log; cut -c1-200 gpurun_out/bench_c4.json
python tools/bench_output.py > gpurun_out/output_stage.jsonl 2>&1; python tools/bench_output.py --sources 20000 --no-cpu >> gpurun_out/output_stage.jsonl 2>&1; cut -c1-300 gpurun_out/output_stage.jsonl
python tools/profile_run.py cornell 1024 1024 2 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_run.py cornell 1024 1024 2 > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/plain.log
