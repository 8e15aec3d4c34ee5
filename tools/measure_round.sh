set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_r01.json
python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_err.log; tail -c 400 gpurun_out/bench_r01_reference.json
python tools/profile_run.py cornell 1024 1024 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_run.py cornell 1024 1024 1 > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/plain.log
python tools/run_configs.py c1,c5_100,c5 > gpurun_out/configs_c1c5.jsonl 2>&1; tail -c 1500 gpurun_out/configs_c1c5.jsonl
