"""How many (sub-tile, splat) pairs does the bounding-box test of the blend kernels admit that an exact ellipse-vs-rectangle
test would reject?  200 k Gaussians of the bench scene (median sigma 1.5 px, 1080p), fp64, CPU only.
Result (this container): gaussians 200000 box pairs 1120918 exact-rect pairs 995356 pixel-any pairs 993256  (DESIGN.md section 6)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from log_b200.synthetic import make_camera, make_scene
from oracle import torch_dense as O
n=200000; W,H=1920,1080
cam=make_camera(W,H,dtype=torch.float64,sh_degree=0)
sc=make_scene(n,W,H,1.5,seed=0,sh_degree=0,dtype=torch.float64)
p=O.project(sc['means3D'],sc['scales'],sc['rotations'],cam,1)
xy=p['xy'].numpy(); a,b,c=[t.numpy() for t in p['cov']]; con=p['conic'].numpy(); valid=p['valid'].numpy()
o=sc['opacities'].numpy().reshape(-1)
ok=valid&(o*255>=1)
xy=xy[ok];a=a[ok];b=b[ok];c=c[ok];con=con[ok];o=o[ok]
q=2*np.log(o*255)
hx=np.sqrt(q*a);hy=np.sqrt(q*c)
box=0;exact=0;pix_any=0; tiles=0
# iterate over subtiles overlapped by box
for i in range(len(o)):
    sx0=int(np.floor((xy[i,0]-hx[i])/8)); sx1=int(np.floor((xy[i,0]+hx[i])/8))
    sy0=int(np.floor((xy[i,1]-hy[i])/4)); sy1=int(np.floor((xy[i,1]+hy[i])/4))
    for sy in range(max(sy0,0),min(sy1,H//4)+1):
        for sx in range(max(sx0,0),min(sx1,W//8-1)+1):
            x0=sx*8;y0=sy*4
            # box test with pixel-centre rectangle [x0,x0+7]x[y0,y0+3]
            if not (xy[i,0]+hx[i]>=x0 and xy[i,0]-hx[i]<=x0+7 and xy[i,1]+hy[i]>=y0 and xy[i,1]-hy[i]<=y0+3): continue
            box+=1
            px=np.arange(x0,x0+8)[None,:]-xy[i,0]; py=np.arange(y0,y0+4)[:,None]-xy[i,1]
            qq=con[i,0]*px*px+2*con[i,1]*px*py+con[i,2]*py*py
            if (qq<=q[i]).any(): pix_any+=1
            # exact continuous rect min
            cx=np.clip(xy[i,0],x0,x0+7);cy=np.clip(xy[i,1],y0,y0+3)
            if cx==xy[i,0] and cy==xy[i,1]: exact+=1; continue
            best=1e30
            A,B,C=con[i,0],con[i,1],con[i,2]
            for (fx,vx) in ((True,x0-xy[i,0]),(True,x0+7-xy[i,0])):
                # x fixed = vx, minimise over dy in [y0-y, y0+3-y]: q= A vx^2+2B vx dy+C dy^2 -> dy*=-B vx/C
                dy=np.clip(-B*vx/C,y0-xy[i,1],y0+3-xy[i,1]); best=min(best,A*vx*vx+2*B*vx*dy+C*dy*dy)
            for vy in (y0-xy[i,1],y0+3-xy[i,1]):
                dx=np.clip(-B*vy/A,x0-xy[i,0],x0+7-xy[i,0]); best=min(best,A*dx*dx+2*B*dx*vy+C*vy*vy)
            if best<=q[i]: exact+=1
print('gaussians',len(o),'box pairs',box,'exact-rect pairs',exact,'pixel-any pairs',pix_any)
