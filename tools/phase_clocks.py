import sys, ctypes as C, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from sdf_tools_b200 import Context, synth, _capi
ctx=Context(0); lib=_capi.load()
m=synth.make_mask("c3"); nx,ny,nz=m.shape
dm=torch.from_numpy(m).cuda(); out=torch.empty(m.shape,dtype=torch.float32,device='cuda')
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(2): ctx.build_device(dm.data_ptr(),nx,ny,nz,0.01,out.data_ptr())
torch.cuda.synchronize()
lib.sdf_b200_debug_phase_clocks(ctx._h,1,None,0)
ctx.build_device(dm.data_ptr(),nx,ny,nz,0.01,out.data_ptr()); torch.cuda.synchronize()
buf=(C.c_longlong*(2*1024*8))()
lib.sdf_b200_debug_phase_clocks(ctx._h,1,buf,2*1024*8)
a=np.array(buf[:],dtype=np.int64).reshape(2,1024,8)
for p in range(2):
    t=a[p][:, :6]; d=np.diff(t,axis=1)
    print("pass",p+2,"median cycles: load %d build %d wait-barrier %d merge %d query %d total %d"%tuple(list(np.median(d,axis=0))+[np.median(t[:,5]-t[:,0])]))
    print("   p90: ", np.percentile(d,90,axis=0).astype(int), " first-wave (CTA<296) medians:", np.median(d[:296],axis=0).astype(int))
