"""Gradient sensitivity of the REFERENCE algorithm (CPU oracle) to TF32-sized conv output noise.

Adds zero-mean noise of relative RMS 5e-4 / 1e-4 to every conv output of the fp32 oracle and reports the cosine between
the noisy and clean parameter gradients of the G step (shapes.yaml, B=2).  Result in the build container: median cosine
0.970 (worst 0.90) at 5e-4, 0.998 (worst 0.989) at 1e-4 - i.e. the softargmax -> warp chain itself amplifies
TF32-level rounding into few-percent gradient direction changes; the tensor-core path measures median 0.985 / worst 0.93
against the fp32 oracle (tests/test_gpu_3_tc.py), inside that envelope."""
import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers
from oracle import monkey_oracle as mo
cfg = helpers.load_config('shapes'); tp = cfg['train_params']
def build():
    torch.manual_seed(0)
    g,d,k = mo.build_from_config(cfg)
    for m in (g,d,k):
        for n,p in m.named_parameters():
            if p.dim()>1: torch.nn.init.kaiming_uniform_(p, a=5**0.5)
            elif 'norm' in n and n.endswith('weight'): torch.nn.init.ones_(p)
            else: torch.nn.init.zeros_(p)
    helpers.perturb_flow_head(g)
    return g,d,k
x = {'source': helpers.smooth_frames(2,1,64,5), 'video': helpers.smooth_frames(2,1,64,6)}
def grads(noise):
    g,d,k = build()
    hooks=[]
    if noise:
        gen = torch.Generator().manual_seed(7)
        def hook(m, inp, out):
            return out + noise*out.detach().pow(2).mean().sqrt()*torch.randn(out.shape, generator=gen)
        for mod in (g,d,k):
            for m in mod.modules():
                if isinstance(m, mo._Conv): hooks.append(m.register_forward_hook(hook))
    for m in (g,d,k): m.train()
    out = mo.generator_full(k,g,d,tp,x)
    sum(v.mean() for v in out[:-2]).backward()
    return {('G.'+n if mod is g else 'K.'+n): p.grad.clone() for mod in (g,k) for n,p in mod.named_parameters() if p.grad is not None}
a = grads(0)
for noise in (5e-4, 1e-4):
    b = grads(noise)
    cs = sorted((float(torch.dot(a[n].flatten(), b[n].flatten())/(a[n].norm()*b[n].norm()+1e-30)), n) for n in a if not helpers.structurally_zero_grad(n[2:]))
    print('noise', noise, 'median', cs[len(cs)//2][0], 'worst', cs[:4])
