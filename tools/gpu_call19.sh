set -x
mkdir -p gpurun_out
timeout 900 python tools/ab_r02.py configs base,m_r2,m_r8,m_s8,m_s32,m_l4,m_l10,m_l16,m_b5 c3,c3_tree,c4 > gpurun_out/ab19_mesh_sweep.log 2>&1; cat gpurun_out/ab19_mesh_sweep.log | cut -c1-110
