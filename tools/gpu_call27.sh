set -x
timeout 600 python tools/ab_r02.py run base,cu2,cu4 c2,c1 > gpurun_out/ab27_child_unroll.log 2>&1; cat gpurun_out/ab27_child_unroll.log
timeout 300 python tools/ab_r02.py configs base,cu2 c5 > gpurun_out/ab27_c5.log 2>&1; cat gpurun_out/ab27_c5.log | cut -c1-100
