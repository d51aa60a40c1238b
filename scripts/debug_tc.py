import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
def run(case, x, w, b, path):
    g = pkg.ConvGeom(*case)
    n,h,wd,c,f,rh,rw,ph,pw,sh,sw,dh,dw = case
    oh = (h - rh - (rh-1)*dh + 2*ph)//sh + 1; ow = (wd - rw - (rw-1)*dw + 2*pw)//sw + 1
    y = torch.full((n*oh*ow*f,), float('nan'), device='cuda')
    ctx.set_conv_path(path)
    ctx.conv_forward(g, x, w, b, y)
    torch.cuda.synchronize()
    return y.cpu().numpy().reshape((n,oh,ow,f), order='F')
case = (32, 4, 4, 32, 32, 1, 1, 0, 0, 1, 1, 0, 0)
n,h,wd,c,f = case[:5]
torch.manual_seed(0)
x = torch.ones(n*h*wd*c, device='cuda'); w = torch.ones(c*f, device='cuda'); b = torch.zeros(f, device='cuda')
yt = run(case, x, w, b, 2); ys = run(case, x, w, b, 1)
print("ones: tc", yt.ravel()[:8], "nan", np.isnan(yt).sum(), "simt", ys.ravel()[:4])
# x varies with n only
xn = torch.arange(n, device='cuda', dtype=torch.float32).repeat(h*wd*c)
yt = run(case, xn, w, b, 2); ys = run(case, xn, w, b, 1)
print("x=n: tc", yt[:6,0,0,0], "simt", ys[:6,0,0,0])
# x varies with c only: x[n,h,w,c] = c ; w = delta(c==k) for filter k
xc = torch.arange(c, device='cuda', dtype=torch.float32).repeat_interleave(n*h*wd)
wk = torch.eye(c, f, device='cuda').t().contiguous().t().reshape(-1)  # W[c + C*f] = (c==f)
wk = torch.zeros(c*f, device='cuda'); 
for i in range(min(c,f)): wk[i + c*i] = 1
yt = run(case, xc, wk, b, 2); ys = run(case, xc, wk, b, 1)
print("x=c,w=I: tc", yt[0,0,0,:8], "simt", ys[0,0,0,:8])
xr = torch.rand(n*h*wd*c, device='cuda'); wr = torch.rand(c*f, device='cuda')
yt = run(case, xr, wr, b, 2); ys = run(case, xr, wr, b, 1)
print("rand maxrel", np.abs(yt-ys).max()/np.abs(ys).max())
