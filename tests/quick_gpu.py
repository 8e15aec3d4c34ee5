import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from ipt_b200 import capi
lib=capi.load()
print('devices', lib.ipt_device_count())
for name,W,H,passes,batch in [('box',640,640,4,0),('box',640,640,4,1<<16),('box',640,640,4,1<<20),('cornell',1024,1024,4,0),('cornell',1024,1024,4,1<<16), ('cornell',1024,1024,4,1<<20)]:
    sd=capi.SceneDescription(name); sc=capi.Scene(sd)
    pl=capi.Plane(sc,W,H)
    p=capi.default_params(width=W,height=H,pass_count=passes,batch_paths=batch)
    pl.render(p)
    p.flags=capi.FLAG_TIME_KERNELS
    st=pl.render(p)
    p.flags=0
    st2=pl.render(p)
    print(name,W,H,'batch',batch,'ms',round(st2.ms_total,2),'Mpaths/s',round(st2.paths/st2.ms_total/1e3,2),'Mrays/s',round(st2.rays/st2.ms_total/1e3,1),'rays/path',round(st2.rays/st2.paths,1),
          'gen %.2f ext %.2f shade %.2f acc %.2f (timed total %.2f)'%(st.ms_generate,st.ms_extend,st.ms_shade,st.ms_accumulate,st.ms_total),'launches',st2.kernel_launches, 'queueGB/s', round(st2.queue_bytes/st2.ms_total/1e6,1))
    print('   depth rays',list(st2.rays_at_depth)[:4],'surf',st2.surface_hits,'light',st2.light_hits,'miss',st2.misses,'failed',st2.failed_samples,'pruned',st2.zero_weight_pruned,'dropped',st2.nonfinite_dropped)
    pl.close(); sc.close()
