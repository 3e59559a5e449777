import sys, ctypes, torch
sys.path.insert(0, ".")
from primia_b200._lib import call, ptr, stream
B,H,W=2,40,40
x=torch.randn(B,3,H,W,device="cuda"); w=torch.randn(64,7,7,3,device="cuda")*0.1
w192=torch.empty(64,192,dtype=torch.bfloat16,device="cuda")
call("pm_stem_prep_w_bf16", ptr(w), ptr(w192), stream())
Ho=Wo=20
y=torch.empty(B,Ho,Wo,64,dtype=torch.bfloat16,device="cuda")
st=torch.zeros(128,dtype=torch.float64,device="cuda")
call("pm_stem_conv_fwd_bf16", ptr(x), ptr(w192), B,H,W, ptr(y), ptr(st), stream())
torch.cuda.synchronize()
print("fwd ok", y.float().abs().mean().item())
