import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import fbstab_b200 as fb, torch
nz,nl,nv,B=32,8,64,int(sys.argv[1]) if len(sys.argv)>1 else 16384
d = fb.problems.random_dense_qp(nz,nl,nv,count=B,config=2)
s = fb.FBstabDense(nz,nl,nv,max_batch=B)
dev=torch.device('cuda:0')
dd={k:torch.from_numpy(a).to(dev) for k,a in d.items()}
for it in range(2):
    zt=torch.zeros(B*nz,dtype=torch.float64,device=dev); lt=torch.zeros(B*nl,dtype=torch.float64,device=dev); vt=torch.zeros(B*nv,dtype=torch.float64,device=dev)
    out,y=s.solve_batch(dd,zt,lt,vt)
    torch.cuda.synchronize()
