import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import oracle_lib
from ipt_b200 import capi
orc=oracle_lib.load_oracle()
sd=capi.SceneDescription('mesh:20000'); sc=capi.Scene(sd)
p=capi.default_params(width=64,height=64,pass_count=1,depth_max=2,schedule=[1,1],flags=2)
s,q,c,st=sc.render_host(p)
o=orc.render(sd.ptr,p,oracle_lib.RNG_PHILOX,1)
bad=np.argwhere(np.abs(s-o['sum'])>1e-4)
print(len(bad), bad[:5])
for (y,x) in bad[:2]:
    # find loop pixel: plane GRID maps loop row iy -> 62-iy ; use LINEAR to be sure
    pass
p=capi.default_params(width=64,height=64,pass_count=1,depth_max=2,schedule=[1,1],flags=2,plane_mode=capi.PLANE_LINEAR)
s,q,c,st=sc.render_host(p)
o=orc.render(sd.ptr,p,oracle_lib.RNG_PHILOX,1)
bad=np.argwhere(np.abs(s-o['sum'])>1e-4)
for (y,x) in bad[:2]:
    pp=capi.default_params(width=64,height=64,pass_count=1,depth_max=2,schedule=[1,1],flags=2|4,plane_mode=capi.PLANE_LINEAR,tile_x0=int(x),tile_y0=int(y),tile_w=1,tile_h=1)
    s1,_,_,_=sc.render_host(pp); sys.stdout.flush()
    o1=orc.render(sd.ptr,pp,oracle_lib.RNG_PHILOX,1)
    print('pixel',y,x,'gpu',s1[y,x],'cpu',o1['sum'][y,x]); sys.stdout.flush()
