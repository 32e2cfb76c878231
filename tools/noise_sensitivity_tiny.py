"""Sensitivity of the ORACLE's own G-step gradients (tiny architecture, the inputs of
tests/test_gpu_2_modules.py::test_train_step_gradient_parity_tiny) to relative noise on its conv outputs.  Build-container
result: noise 1e-7 -> median 7e-6 / worst 1e-4; 1e-6 -> median 7e-3 / worst 2e-1; 1e-5 -> median 2e-2 / worst 2e-1, for every
input seed tried: the composed gradient is ill-conditioned in the reference algorithm itself."""
import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers
from oracle import monkey_oracle as mo
import test_gpu_2_modules as t2
cfg = helpers.tiny_config(); tp = cfg['train_params']
gen, disc, kp = t2.build_product(cfg)
sd = [m.state_dict() for m in (gen, disc, kp)]
x = {'source': helpers.smooth_frames(3,1,32,5), 'video': helpers.smooth_frames(3,1,32,6)}
def grads(noise, seed=7):
    g,d,k = mo.build_from_config(cfg)
    g.load_state_dict(sd[0]); d.load_state_dict(sd[1]); k.load_state_dict(sd[2])
    if noise:
        rng = torch.Generator().manual_seed(seed)
        def hook(m, inp, out):
            return out + noise*out.detach().pow(2).mean().sqrt()*torch.randn(out.shape, generator=rng)
        for mod in (g,d,k):
            for m in mod.modules():
                if isinstance(m, mo._Conv): m.register_forward_hook(hook)
    for m in (g,d,k): m.train()
    out = mo.generator_full(k,g,d,tp,x)
    sum(v.mean() for v in out[:-2]).backward()
    return {n: p.grad.clone() for mod in (g,k,d) for n,p in mod.named_parameters() if p.grad is not None}
a = grads(0)
for noise in (1e-7, 1e-6, 1e-5):
    b = grads(noise)
    es = sorted(((helpers.rel_err(b[n], a[n]), n) for n in a if not helpers.structurally_zero_grad(n)), reverse=True)
    print('noise', noise, 'median %.2e' % es[len(es)//2][0], 'worst', [(round(e,5), n) for e,n in es[:3]])
print('--- seed scan')
for s1 in (11, 21, 31, 41, 51):
    x = {'source': helpers.smooth_frames(3,1,32,s1), 'video': helpers.smooth_frames(3,1,32,s1+1)}
    a = grads(0)
    res=[]
    for noise in (1e-6, 1e-5):
        b = grads(noise)
        es = sorted(((helpers.rel_err(b[n], a[n]), n) for n in a if not helpers.structurally_zero_grad(n)), reverse=True)
        res.append((noise, 'med %.1e worst %.1e' % (es[len(es)//2][0], es[0][0])))
    print(s1, res)
