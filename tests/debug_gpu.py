import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import oracle_lib
from ipt_b200 import capi
orc=oracle_lib.load_oracle()
sd=capi.SceneDescription('box'); sc=capi.Scene(sd)
for (y,x) in [(35,78),(56,69)]:
    p=capi.default_params(width=96,height=96,pass_count=1,depth_max=3,schedule=[1,1,1],flags=capi.FLAG_KEEP_ZERO_WEIGHT|4,plane_mode=capi.PLANE_LINEAR,tile_x0=x,tile_y0=y,tile_w=1,tile_h=1)
    s,q,c,st=sc.render_host(p)
    sys.stdout.flush()
    o=orc.render(sd.ptr,p,oracle_lib.RNG_PHILOX,0)
    print('pixel',y,x,'gpu',s[y,x],'cpu',o['sum'][y,x]); sys.stdout.flush()
