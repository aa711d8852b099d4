import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import fbstab_b200 as fb
from oracle import binding as ob
nz,nl,nv=32,8,64; B=8
d=fb.problems.random_dense_qp(nz,nl,nv,count=B,config=2)
sz={"H":nz*nz,"f":nz,"G":nl*nz,"h":nl,"A":nv*nz,"b":nv}
s=fb.FBstabDense(nz,nl,nv,max_batch=B)
for sigma in (1e-3,1e-8):
  for which in ('mid','sol'):
    Z=[];L=[];V=[];Y=[];RZ=[];RL=[];RV=[];ref0=[];ref3=[]
    for i in range(B):
        p=ob.Problem.dense(*[d[k][i*sz[k]:(i+1)*sz[k]] for k in fb.problems.DENSE_FIELDS])
        o=ob.default_options()
        if which=='mid': o.max_newton_iters=6
        out,(z,l,v,y),_=p.solve(o)
        rz,rl,rv,_=p.residual('inner',(z,l,v,y),(z,l,v),sigma=sigma)
        Z.append(z);L.append(l);V.append(v);Y.append(y);RZ.append(rz);RL.append(rl);RV.append(rv)
        for var,ref in ((0,ref0),(3,ref3)):
            rc,dx,_,_=p.linear_solve((z,l,v,y),(z,l,v),sigma,(rz,rl,rv),variant=var)
            ref.append(np.concatenate(dx))
    cat=lambda a: np.ascontiguousarray(np.concatenate(a))
    z,l,v,y,rz,rl,rv=map(cat,(Z,L,V,Y,RZ,RL,RV))
    dz,dl,dv,dy=np.zeros(B*nz),np.zeros(B*nl),np.zeros(B*nv),np.zeros(B*nv)
    st=np.zeros(B,dtype=np.int32)
    s.component(fb.capi.COMP_NEWTON,d,B,z=z,l=l,v=v,y=y,zbar=z,lbar=l,vbar=v,rz=rz.copy(),rl=rl.copy(),rv=rv.copy(),dz=dz,dl=dl,dv=dv,dy=dy,status=st,sigma=sigma)
    for i in range(B):
        g=np.concatenate([dz[i*nz:(i+1)*nz],dl[i*nl:(i+1)*nl],dv[i*nv:(i+1)*nv],dy[i*nv:(i+1)*nv]])
        n=np.abs(ref0[i]).max()
        print(sigma,which,i,'gpu-vs-ldlt %.2e'%(np.abs(g-ref0[i]).max()/n),'gj-vs-ldlt %.2e'%(np.abs(ref3[i]-ref0[i]).max()/n),'parts z %.1e l %.1e v %.1e'%(np.abs(g[:nz]-ref0[i][:nz]).max()/n,np.abs(g[nz:nz+nl]-ref0[i][nz:nz+nl]).max()/n,np.abs(g[nz+nl:nz+nl+nv]-ref0[i][nz+nl:nz+nl+nv]).max()/n), st[i])
