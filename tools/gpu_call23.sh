set -x
timeout 900 python tools/ab_r02.py configs base,m_v2,m_v2s16,m_v3 c3,c3_tree,c4 > gpurun_out/ab23_mesh_visits.log 2>&1; cat gpurun_out/ab23_mesh_visits.log | cut -c1-100
