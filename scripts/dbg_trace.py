import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo')
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
ctx=floor_b200.device_context(); dev=ctx.get_device(0); q=ctx.create_queue(dev)
M=T.FLAG_MIPMAPPED|T.READ_WRITE
dim,t=((8192,8192),T.IMAGE_2D|T.RGBA16F|M)
img=ctx.create_image(q,dim,t); img.fill_synthetic(q,2); q.finish()
img.enqueue_mip_map_chain(q); q.finish()
L=floor_b200.lib(); L.flmip_debug_read.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
N=32+2*60000
out=(ctypes.c_uint64*N)()
L.flmip_debug_read(img._handle, out, N)
a=np.frombuffer(out,dtype=np.uint64)
n=int(a[16]); print('events',n)
ev=a[32:32+2*n].reshape(-1,2)
typ=(ev[:,0]>>np.uint64(48)).astype(int); extra=((ev[:,0]>>np.uint64(24))&np.uint64(0xFFFFFF)).astype(int); cta=(ev[:,0]&np.uint64(0xFFFFFF)).astype(int); ts=ev[:,1].astype(np.int64)
t0=ts[typ==1].min()
print('start spread', ts[typ==1].max()-t0)
end=ts[typ==2]-t0; print('consumer end: min',end.min(),'median',np.median(end),'max',end.max())

g0=ts[typ==3]-t0; g1=ts[typ==4]-t0
print('group stages',len(g0),'begin times',np.sort(g0)[:70:4], 'end max', g1.max())
# progress: time to reach iteration k per cta
for k in (8,16,24,32,40,48):
    sel=(typ==5)&(extra==k)
    if sel.any(): print('iter',k,'time min/med/max', (ts[sel]-t0).min(), np.median(ts[sel]-t0), (ts[sel]-t0).max())
img.destroy()
# group-stage timelines of the last 4 groups (by begin time)
import collections
order=np.argsort(ts)
per=collections.defaultdict(list)
for i in order:
    if typ[i] in (3,7,8,9,10,11,4): per[cta[i]].append((typ[i], int(ts[i]-t0)))
last=sorted(per.items(), key=lambda kv: -max(x[1] for x in kv[1]))[:4]
for c,evs in last: print('cta',c,evs[-14:])
# per-tile finisher durations
d={}
for i in order:
    if typ[i] in (12,13,14): d.setdefault((cta[i],extra[i]),{})[typ[i]]=int(ts[i])
casc=[v[13]-v[12] for v in d.values() if 12 in v and 13 in v]; pub=[v[14]-v[13] for v in d.values() if 13 in v and 14 in v]
print('tile cascade ns: median',np.median(casc),'p90',np.percentile(casc,90),'max',max(casc)); print('tile publish ns: median',np.median(pub),'p90',np.percentile(pub,90),'max',max(pub))
