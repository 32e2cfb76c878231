"""Diagnostic: taichi.yaml @256, B=2, train-mode G-step: exact-FFMA and 3xTF32 product paths against the CPU oracle and
against each other - separates kernel error from the conditioning of the train-mode network itself."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import helpers  # noqa: E402
import train_glue  # noqa: E402
from oracle import monkey_oracle as mo  # noqa: E402
from monkey_net_b200 import ops  # noqa: E402
import test_gpu_2_modules as t2  # noqa: E402


def run(mode, cfg, x):
    (gen, disc, kp), _, _ = t2._pair(cfg, 256, 2)
    for m in (gen, disc, kp):
        m.train()
    ops.set_conv_mode(mode)
    out = train_glue.generator_full(kp, gen, disc, cfg['train_params'], {k: v.cuda() for k, v in x.items()})
    sum(v.mean() for v in out[:-2]).backward()
    torch.cuda.synchronize()
    ops.set_conv_mode('auto')
    grads = {n: p.grad.detach().cpu() for m in (gen, kp) for n, p in m.named_parameters() if p.grad is not None}
    return out[-2]['video_prediction'].detach().cpu(), out[-1]['mean'].detach().cpu(), [float(v.mean()) for v in out[:-2]], grads


def cmp(tag, a, b):
    fa, ka, la, ga = a
    fb, kb, lb, gb = b
    d = (fa - fb).abs()
    coss = []
    for n in ga:
        if n in gb and not helpers.structurally_zero_grad(n):
            u, v = ga[n].flatten(), gb[n].flatten()
            coss.append(float(torch.dot(u, v) / (u.norm() * v.norm() + 1e-30)))
    coss.sort()
    print('%-22s frame max %.2e mean %.2e  >1e-3: %d px  kp %.2e  loss rel %.2e  cos med %.6f p10 %.6f min %.4f'
          % (tag, float(d.max()), float(d.mean()), int((d > 1e-3).sum()), float((ka - kb).abs().max()),
             max(abs(p - q) / max(abs(q), 1e-9) for p, q in zip(la, lb)), coss[len(coss) // 2], coss[len(coss) // 10],
             coss[0]), flush=True)


def main():
    cfg = helpers.load_config('taichi')
    (gen, disc, kp), (og, od, ok), x = t2._pair(cfg, 256, 2)
    for m in (og, od, ok):
        m.train()
    out = mo.generator_full(ok, og, od, cfg['train_params'], x)
    sum(v.mean() for v in out[:-2]).backward()
    ref = (out[-2]['video_prediction'].detach(), out[-1]['mean'].detach(), [float(v.mean()) for v in out[:-2]],
           {n: p.grad for m in (og, ok) for n, p in m.named_parameters() if p.grad is not None})
    del gen, disc, kp
    a = run('fp32', cfg, x)
    b = run('tf32x3', cfg, x)
    a2 = run('fp32', cfg, x)
    cmp('fp32 vs oracle', a, ref)
    cmp('tf32x3 vs oracle', b, ref)
    cmp('tf32x3 vs fp32', b, a)
    cmp('fp32 vs fp32 (rerun)', a2, a)


if __name__ == '__main__':
    main()
