import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo')
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
ctx=floor_b200.device_context(); dev=ctx.get_device(0); q=ctx.create_queue(dev)
M=T.FLAG_MIPMAPPED|T.READ_WRITE
for dim,t in [((8192,8192),T.IMAGE_2D|T.RGBA16F|M), ((4096,4096,1),T.IMAGE_CUBE_ARRAY|T.RGBA32F|M)]:
    img=ctx.create_image(q,dim,t); img.fill_synthetic(q,2); q.finish()
    for i in range(3): img.enqueue_mip_map_chain(q)
    q.finish()
    L=floor_b200.lib(); L.flmip_debug_read.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    out=(ctypes.c_uint64*16)()
    L.flmip_debug_read(img._handle, out, 16)
    o=list(out)
    n=max(o[0],1); m=max(o[6],1)
    print(dim, 'groups',o[0],'avg ns: lock',o[1]/n,'gather',o[2]/n,'cascade',o[3]/n,'layer+unlock',o[4]/n,'max total',o[5],'| publishes',o[6],'avg',o[7]/m,'max',o[8])
    img.destroy()
