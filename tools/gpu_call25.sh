set -x
timeout 900 python tools/ab_r02.py configs base,m_l4,m_l8,m_l10,m_r2,m_r8 c3,c4 > gpurun_out/ab25_mesh_retune.log 2>&1; cat gpurun_out/ab25_mesh_retune.log | cut -c1-100
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest25.log 2>&1; tail -3 gpurun_out/pytest25.log
