"""BASELINE.json configs[0] (the reference's own CPU-runnable case): config/shapes.yaml as shipped, eval mode, B=1,
reconstruction of the bundled 32-frame data/shapes test video - the product path (one KP pass over the 32 frames + one
generator pass, fused inference kernels) against the fixture recorded from the reference's own `generate`
(oracle/make_golden.py:run_reconstruction).  Keypoints <= 2e-5 and identical pixel indices, frames <= 1e-3."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


# measured on a B200 (round 2): fp32 |kp| 3.5e-8 |frame| 4.8e-7; tf32 |kp| 2.6e-5 |frame| 3.8e-4.  'auto' is the
# product default (geometry networks tf32x3, appearance path 1xTF32): the north-star bars incl. identical pixel indices
@pytest.mark.parametrize('mode,tol_kp,tol_frame', [('fp32', 2e-5, 1e-4), ('auto', 2e-5, 1e-3), ('tf32', 1e-4, 1e-3)])
def test_reconstruction_of_bundled_shapes_video(mode, tol_kp, tol_frame):
    from monkey_net_b200 import ops, transfer_step
    import test_gpu_2_modules as t2
    import test_oracle_golden as tog
    gold = helpers.load_golden('golden_recon_shapes')
    cfg = helpers.load_config('shapes')
    gen, disc, kp = t2.build_product(cfg)
    helpers.assert_checksums([helpers.state_checksum(gen.state_dict()), helpers.state_checksum(kp.state_dict())],
                             gold['checksum'])
    for m in (gen, kp):
        m.cuda().eval()
    video = tog.recon_inputs(gold).cuda()
    prev = ops.CONV_MODE
    ops.set_conv_mode(mode)
    try:
        with torch.no_grad():
            out = transfer_step.reconstruct(gen, kp, video)
            loop = transfer_step.reconstruct(gen, kp, video[:, :, :3], batched=False)
    finally:
        ops.set_conv_mode(prev)
    keep = gold['keep'].tolist()
    ref_mean = torch.from_numpy(gold['kp_mean'])
    print('reconstruction [%s]: |kp| %.2e  |frame| %.2e' % (
        mode, helpers.max_abs(out['kp_driving']['mean'], ref_mean),
        helpers.max_abs(out['video_prediction'][:, :, keep], torch.from_numpy(gold['video_prediction']))))
    assert out['video_prediction'].shape == (1, 3, 32, 64, 64)
    assert helpers.max_abs(out['kp_driving']['mean'], ref_mean) < tol_kp
    assert helpers.max_abs(out['video_prediction'][:, :, keep], torch.from_numpy(gold['video_prediction'])) < tol_frame
    assert helpers.max_abs(loop['video_prediction'], out['video_prediction'][:, :, :3]) < tol_frame
    if mode in ('fp32', 'auto'):
        assert helpers.max_abs(out['kp_driving']['var'], torch.from_numpy(gold['kp_var'])) < tol_kp
        assert helpers.max_abs(out['video_deformed'][:, :, keep], torch.from_numpy(gold['video_deformed'])) < 1e-3
        # "keypoint indices bit-exact": the pixel the visualiser draws (logger.py:99-100) is identical
        px = lambda m: torch.round(64 * (m + 1) / 2)
        assert torch.equal(px(out['kp_driving']['mean'].cpu()), px(ref_mean))
