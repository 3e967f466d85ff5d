import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O
import multimodal_pl_b200 as mm
ops = mm.ops
mm.set_compute_dtype(torch.float32); mm.set_conv_algo("direct")
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def R(shape, seed, scale=1.0): return scale * torch.randn(shape, generator=torch.Generator().manual_seed(seed))

for shape in [(1,128,4,8,8), (1,64,8,16,16), (1,256,2,4,4), (1,32,16,32,32)]:
    C = shape[1]
    x = R(shape,1)+0.3; g1,b1,g2,b2 = 1+0.2*R((C,),2), 0.2*R((C,),3), 1+0.2*R((C,),4), 0.2*R((C,),5)
    dy1, dy2 = R(shape,6), R(shape,7)
    xr = x.double().requires_grad_(True); pr=[t.double().requires_grad_(True) for t in (g1,b1,g2,b2)]
    y1r, y2r = O.gn_relu(xr,pr[0],pr[1]), O.gn_relu(xr,pr[2],pr[3])
    ((y1r*dy1.double()).sum()+(y2r*dy2.double()).sum()).backward()
    x32 = x.clone().requires_grad_(True); p32=[t.clone().requires_grad_(True) for t in (g1,b1,g2,b2)]
    ((O.gn_relu(x32,p32[0],p32[1])*dy1).sum()+(O.gn_relu(x32,p32[2],p32[3])*dy2).sum()).backward()
    xd = x.cuda().requires_grad_(True); pd=[t.cuda().requires_grad_(True) for t in (g1,b1,g2,b2)]
    a,b = ops.gn_relu_dual(xd,*pd)
    ((a*dy1.cuda()).sum()+(b*dy2.cuda()).sum()).backward()
    print(f"GN dual {shape}: fwd {rel(a,y1r):.2e} | dx mine {rel(xd.grad,xr.grad):.2e} torch32 {rel(x32.grad,xr.grad):.2e} | "
          + " ".join(f"{n} {rel(pd[i].grad,pr[i].grad):.1e}/{rel(p32[i].grad,pr[i].grad):.1e}" for i,n in enumerate(["dg1","db1","dg2","db2"])))

for (cin,cout,k,stride,sp) in [(128,64,3,1,(4,8,8)), (128,64,1,1,(4,8,8)), (64,128,3,2,(8,16,16)), (32,32,3,1,(16,32,32))]:
    x = R((1,cin)+sp,1); w = R((cout,cin,k,k,k),2)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = O.ws_conv3d(xr,wr,stride,k//2); dy = R(tuple(yr.shape),3); (yr*dy.double()).sum().backward()
    x32, w32 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y32 = O.ws_conv3d(x32,w32,stride,k//2); (y32*dy).sum().backward()
    xd, wd = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    y = ops.ws_conv3d(xd,wd,stride,True,None); y.backward(dy.cuda())
    print(f"conv {cin}->{cout} k{k} s{stride} {sp}: y mine {rel(y,yr):.2e} t32 {rel(y32,yr):.2e} | dx mine {rel(xd.grad,xr.grad):.2e} t32 {rel(x32.grad,xr.grad):.2e} | dw mine {rel(wd.grad,wr.grad):.2e} t32 {rel(w32.grad,wr.grad):.2e}")

for shape in [(1,128,4,8,8),(1,256,2,4,4)]:
    n,c,d,h,w = shape
    x = R(shape,1); skip = R((n,c,2*d,2*h,2*w),2); dy = R(skip.shape,3)
    xr, sr = x.double().requires_grad_(True), skip.double().requires_grad_(True)
    (O.upsample2x_add(xr,sr)*dy.double()).sum().backward()
    xd, sd = x.cuda().requires_grad_(True), skip.cuda().requires_grad_(True)
    y = ops.upsample2x_add(xd,sd); y.backward(dy.cuda())
    print(f"upsample {shape}: dx {rel(xd.grad,xr.grad):.2e}")
