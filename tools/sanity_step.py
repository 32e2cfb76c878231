"""Run-to-run consistency of one G-step (forward + backward) on the tiny architecture: the same computation twice on
the GPU (differences = summation-order nondeterminism of the atomics) and against the CPU oracle.  Also the target of
`compute-sanitizer --tool {memcheck,initcheck,racecheck}` runs (small enough to finish under the sanitizer).

    python tools/sanity_step.py [fp32|tf32] [--no-oracle]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import helpers  # noqa: E402
import train_glue  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 'fp32'
    from monkey_net_b200 import ops
    import test_gpu_2_modules as t2
    from oracle import monkey_oracle as mo
    ops.set_conv_mode(mode)
    cfg = helpers.tiny_config()
    (gen, disc, kp), (og, od, ok), x = t2._pair(cfg, 32, 3)
    tp = cfg['train_params']
    for m in (gen, disc, kp, og, od, ok):
        m.train()
    xg = {k: v.cuda() for k, v in x.items()}
    runs = []
    for r in range(2):
        for m in (gen, disc, kp):
            m.zero_grad()
        out = train_glue.generator_full(kp, gen, disc, tp, xg)
        sum(v.mean() for v in out[:-2]).backward()
        torch.cuda.synchronize()
        runs.append({n: p.grad.detach().clone() for m in (gen, kp, disc) for n, p in m.named_parameters()
                     if p.grad is not None})
    names = [n for n in runs[0] if not helpers.structurally_zero_grad(n)]
    self_err = sorted(((helpers.rel_err(runs[0][n], runs[1][n]), n) for n in names), reverse=True)
    print('[%s] GPU run 1 vs GPU run 2, worst relative gradient differences:' % mode)
    for e, n in self_err[:5]:
        print('   %.2e  %s' % (e, n))
    if '--no-oracle' in sys.argv:
        return
    ref = mo.generator_full(ok, og, od, tp, x)
    sum(v.mean() for v in ref[:-2]).backward()
    refg = {n: p.grad for m in (og, ok, od) for n, p in m.named_parameters() if p.grad is not None}
    errs = sorted(((helpers.rel_err(runs[1][n], refg[n]), n) for n in names if n in refg), reverse=True)
    print('[%s] GPU vs CPU oracle, worst relative gradient differences (median %.2e):' % (mode, errs[len(errs) // 2][0]))
    for e, n in errs[:5]:
        print('   %.2e  %s' % (e, n))


if __name__ == '__main__':
    main()
