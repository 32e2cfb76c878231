"""GraphedTrainer: one CUDA-graph launch per training iteration must be the same computation as the eager
train.py:110-136 loop body (same losses, same parameter trajectory), and its warm-up must not advance training."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _nets(cfg):
    import test_gpu_2_modules as t2
    gen, disc, kp = t2.build_product(cfg)
    for m in (gen, disc, kp):
        m.cuda().train()
    return gen, disc, kp


def test_graphed_step_equals_eager_step():
    from monkey_net_b200 import train_step
    cfg = helpers.tiny_config()
    tp = cfg['train_params']
    batches = [{'source': helpers.smooth_frames(4, 1, 32, 10 + i).cuda(), 'video': helpers.smooth_frames(4, 1, 32, 20 + i).cuda()}
               for i in range(3)]
    ga, da, ka = _nets(cfg)
    gb, db, kb = _nets(cfg)
    eager = train_step.GraphedTrainer(ka, ga, da, tp, use_graph=False)
    graphed = train_step.GraphedTrainer(kb, gb, db, tp, use_graph=True)
    for x in batches:
        la = eager.step(x).clone()
        lb = graphed.step(x).clone()
        # same kernels, same order; the capturable Adam differs in rounding from the eager one and the ill-conditioned
        # gradients (see test_train_step_gradient_parity_tiny) amplify that over the steps: losses to 2e-3
        assert helpers.max_abs(la, lb) < 2e-3 * max(1.0, float(la.abs().max())), (la, lb)
    assert graphed.graph is not None and graphed.kernels_per_step > 100
    for (n1, p1), (n2, p2) in zip(ga.named_parameters(), gb.named_parameters()):
        if helpers.structurally_zero_grad(n1):
            continue  # Adam turns rounding-noise gradients into +-lr steps in both runs
        # Adam moves a parameter by at most ~lr per step whatever the gradient's size, so two runs whose gradients
        # differ only in the atomics' summation order may differ by a fraction of lr on near-zero-gradient entries:
        # bar = the 3 steps taken at most, and 10 % of a step on average
        diff = (p1.detach() - p2.detach()).abs()
        assert float(diff.max()) < 3 * tp['lr'] and float(diff.mean()) < 0.1 * tp['lr'], (n1, float(diff.max()))
    rm_a = ga.appearance_encoder.down_blocks[0].norm
    rm_b = gb.appearance_encoder.down_blocks[0].norm
    assert int(rm_a.num_batches_tracked) == int(rm_b.num_batches_tracked) == 3
    assert helpers.max_abs(rm_a.running_var, rm_b.running_var) < 1e-5


def test_pinned_host_batch_path():
    from monkey_net_b200 import train_step
    cfg = helpers.tiny_config()
    gen, disc, kp = _nets(cfg)
    tr = train_step.GraphedTrainer(kp, gen, disc, cfg['train_params'], use_graph=True)
    x = {'source': helpers.smooth_frames(2, 1, 32, 1).pin_memory(), 'video': helpers.smooth_frames(2, 1, 32, 2).pin_memory()}
    l1 = tr.step(x).cpu()
    l2 = tr.step(x).cpu()
    assert torch.isfinite(l1).all() and torch.isfinite(l2).all() and not torch.equal(l1, l2)
