import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ipt_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sd = capi.SceneDescription(f"mesh:{n}"); sc = capi.Scene(sd); pl = capi.Plane(sc, W, W)
st = pl.render(capi.default_params(width=W, height=W, pass_count=1, flags=capi.FLAG_TIME_KERNELS))
print('mesh', n, W, 'ms', st.ms_total, 'ext', st.ms_extend, 'rays', st.rays, 'Mrays/s', st.rays / st.ms_total / 1e3, 'nodes/ray', st.bvh_nodes_visited / st.rays, 'tris/ray', st.triangles_tested / st.rays)
