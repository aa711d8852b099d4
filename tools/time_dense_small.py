import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import fbstab_b200 as fb, torch
nz,nl,nv,B=32,8,64,65536
d = fb.problems.random_dense_qp(nz,nl,nv,count=B,config=2)
s = fb.FBstabDense(nz,nl,nv,max_batch=B)
dev=torch.device('cuda:0')
dd={k:torch.from_numpy(a).to(dev) for k,a in d.items()}
for it in range(4):
    zt=torch.zeros(B*nz,dtype=torch.float64,device=dev); lt=torch.zeros(B*nl,dtype=torch.float64,device=dev); vt=torch.zeros(B*nv,dtype=torch.float64,device=dev)
    torch.cuda.synchronize(); t=time.time()
    out,y=s.solve_batch(dd,zt,lt,vt)
    torch.cuda.synchronize(); t=time.time()-t
    o=np.frombuffer(out.cpu().numpy().tobytes(),dtype=fb.OUT_DTYPE)
    print('dense',nz,nl,nv,B,s.path[:20],'%.2f ms'%(t*1e3),'%.0f solves/s'%(B/t), 'flags',np.bincount(o['eflag'],minlength=6),'newton',o['newton_iters'].mean())
