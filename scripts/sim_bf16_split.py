"""CPU simulation: how accurate is a ResNet-18 step whose convolutions multiply bf16-SPLIT operands with exact (fp32/fp64) accumulation?
   python scripts/sim_bf16_split.py 3   # 2-way split, 3 products ("bf16x3"): logits 3.3e-5, worst gradient 1.6e-2
   python scripts/sim_bf16_split.py 6   # 3-way split, 6 products: logits 2.1e-6, worst gradient 5.5e-6 (fp32 level)
Quoted in DESIGN.md section 2 (why the fp32-accurate tensor-core mode was not built)."""
import torch, copy, sys
sys.path.insert(0,'/root/repo')
from oracle import train_oracle as O
import torch.nn.functional as F
torch.set_num_threads(8)
def split(x, n):
    parts=[]; r=x.clone()
    for i in range(n):
        p=r.to(torch.bfloat16).to(torch.float32); parts.append(p); r=r-p
    return parts
TERMS={3:[(0,0),(0,1),(1,0)], 4:[(0,0),(0,1),(1,0),(1,1)], 6:[(0,0),(0,1),(1,0),(0,2),(1,1),(2,0)]}
NSPLIT={3:2,4:2,6:3}
MODE=int(sys.argv[1]); ACC=torch.float64 if len(sys.argv)>2 and sys.argv[2]=='acc64' else torch.float32
class SplitConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx,x,w,stride,pad):
        ctx.save_for_backward(x,w); ctx.s=stride; ctx.p=pad
        xs=split(x,NSPLIT[MODE]); ws=split(w,NSPLIT[MODE])
        y=0
        for i,j in TERMS[MODE]:
            y=y+F.conv2d(xs[i].to(ACC),ws[j].to(ACC),None,stride,pad)
        return y.float()
    @staticmethod
    def backward(ctx,dy):
        x,w=ctx.saved_tensors
        xs=split(x,NSPLIT[MODE]); ws=split(w,NSPLIT[MODE]); ds=split(dy,NSPLIT[MODE])
        dx=0; dw=0
        for i,j in TERMS[MODE]:
            dx=dx+torch.nn.grad.conv2d_input(x.shape,ws[j].to(ACC),ds[i].to(ACC),ctx.s,ctx.p)
            dw=dw+torch.nn.grad.conv2d_weight(xs[i].to(ACC),w.shape,ds[j].to(ACC),ctx.s,ctx.p)
        return dx.float(),dw.float(),None,None
class SConv(torch.nn.Module):
    def __init__(s,c): super().__init__(); s.weight=c.weight; s.stride=c.stride[0]; s.pad=c.padding[0]
    def forward(s,x): return SplitConv.apply(x,s.weight,s.stride,s.pad)
def swap(m):
    for n,c in list(m.named_children()):
        if isinstance(c,torch.nn.Conv2d): setattr(m,n,SConv(c))
        else: swap(c)
B,size=8,int(sys.argv[3]) if len(sys.argv)>3 else 96
torch.manual_seed(42); m=O.ResNet18(input_size=size); m2=copy.deepcopy(m); m64=copy.deepcopy(m).double(); swap(m2)
g=torch.Generator().manual_seed(42); x=torch.randn(B,3,size,size,generator=g); y=torch.randint(0,3,(B,),generator=g)
def run(mm,x):
    mm.train(); out=mm(x); l=F.cross_entropy(out,y); l.backward(); return out.detach(),l.item()
o1,l1=run(m,x); o2,l2=run(m2,x); o64,l64=run(m64,x.double())
rel=lambda a,b:((a.double()-b.double()).norm()/b.double().norm()).item()
print('logits rel x3 vs f32',rel(o2,o1),' f32 vs f64',rel(o1,o64),' x3 vs f64',rel(o2,o64))
print('loss',abs(l2-l1)/l1)
w=0;ws=None
for (n,p),(n2,p2),(n3,p3) in zip(m.named_parameters(),m2.named_parameters(),m64.named_parameters()):
    e=rel(p2.grad,p.grad); e64=rel(p2.grad,p3.grad); c64=rel(p.grad,p3.grad)
    if e>w: w=e;ws=(n,e,e64,c64)
print('worst grad (name, x3 vs f32, x3 vs f64, f32 vs f64)',ws)
