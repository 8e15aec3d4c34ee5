set -x
timeout 900 python tools/ab_r02.py configs m_v2s16,m_v2s32,m_v3s16,m_v4s16,m_v3s32 c3,c3_tree,c4 > gpurun_out/ab24_mesh_visits.log 2>&1; cat gpurun_out/ab24_mesh_visits.log | cut -c1-100
