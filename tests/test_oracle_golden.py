"""Pins the CPU oracle (oracle/monkey_oracle.py) to the reference:
  * against the committed golden fixtures produced by the UNMODIFIED reference (oracle/make_golden.py);
  * against the live reference when /root/reference is present (build container only).
Also checks that the product's `modules/*` reproduce the reference's default initialisation bit-exactly."""
import pytest
import torch

import helpers
from oracle import monkey_oracle as mo, ref_shim


def _oracle_protocol(cfg, gold, weights_from_gold):
    gen, disc, kp = mo.build_from_config(cfg)
    if weights_from_gold:
        for tag, m in (('G', gen), ('D', disc), ('K', kp)):
            m.load_state_dict(helpers.golden_weights(gold, tag))
    else:
        from modules.generator import MotionTransferGenerator
        from modules.discriminator import Discriminator
        from modules.keypoint_detector import KPDetector
        mp = cfg['model_params']
        torch.manual_seed(0)
        pg = MotionTransferGenerator(**mp['generator_params'], **mp['common_params'])
        pd = Discriminator(**mp['discriminator_params'], **mp['common_params'])
        pk = KPDetector(**mp['kp_detector_params'], **mp['common_params'])
        helpers.perturb_flow_head(pg)
        sums = [helpers.state_checksum(m.state_dict()) for m in (pg, pd, pk)]
        helpers.assert_checksums(sums, gold['checksum'])
        gen.load_state_dict(pg.state_dict()); disc.load_state_dict(pd.state_dict()); kp.load_state_dict(pk.state_dict())
    x = {'source': torch.from_numpy(gold['source']), 'video': torch.from_numpy(gold['video'])}
    return helpers.run_protocol(gen, disc, kp, cfg, x, mo.generator_loss, mo.discriminator_loss)


def test_oracle_matches_golden_tiny():
    gold = helpers.load_golden('golden_tiny')
    rep = helpers.compare_with_golden(_oracle_protocol(helpers.tiny_config(), gold, True), gold)
    print(rep)


def test_oracle_matches_golden_shapes_and_init_parity():
    gold = helpers.load_golden('golden_shapes')
    rep = helpers.compare_with_golden(_oracle_protocol(helpers.load_config('shapes'), gold, False), gold)
    print(rep)


def recon_inputs(gold):
    """(1,3,T,H,W) float video of the bundled data/shapes test strip stored in the fixture as uint8 frames."""
    return torch.from_numpy(gold['frames_u8'].astype('float32') / 255.0).permute(3, 0, 1, 2)[None].contiguous()


def test_oracle_reconstruction_matches_reference_on_bundled_shapes_video():
    """BASELINE.json configs[0]: shapes.yaml, eval, B=1, the 32-frame bundled test video, against the fixture made by
    the reference's own `generate` (oracle/make_golden.py:run_reconstruction)."""
    gold = helpers.load_golden('golden_recon_shapes')
    cfg = helpers.load_config('shapes')
    from modules.generator import MotionTransferGenerator
    from modules.discriminator import Discriminator
    from modules.keypoint_detector import KPDetector
    mp = cfg['model_params']
    torch.manual_seed(0)
    pg = MotionTransferGenerator(**mp['generator_params'], **mp['common_params'])
    Discriminator(**mp['discriminator_params'], **mp['common_params'])  # consumes the RNG exactly like run.py:50-63
    pk = KPDetector(**mp['kp_detector_params'], **mp['common_params'])
    helpers.perturb_flow_head(pg)
    helpers.assert_checksums([helpers.state_checksum(pg.state_dict()), helpers.state_checksum(pk.state_dict())],
                             gold['checksum'])
    og, od, ok = mo.build_from_config(cfg)
    og.load_state_dict(pg.state_dict()); ok.load_state_dict(pk.state_dict())
    for m in (og, ok):
        m.eval()
    with torch.no_grad():
        out = mo.reconstruct(og, ok, recon_inputs(gold))
    keep = gold['keep'].tolist()
    assert helpers.max_abs(out['kp_driving']['mean'], torch.from_numpy(gold['kp_mean'])) < 1e-6
    assert helpers.max_abs(out['kp_driving']['var'], torch.from_numpy(gold['kp_var'])) < 1e-6
    assert helpers.max_abs(out['video_prediction'][:, :, keep], torch.from_numpy(gold['video_prediction'])) < 1e-5
    # the warped SOURCE frame has 0 -> 1 edges between neighbouring pixels: 1e-6-pixel coordinate differences of the
    # resized grid (batched vs per-frame evaluation order) show up as 1e-5 intensity differences on edge pixels
    assert helpers.max_abs(out['video_deformed'][:, :, keep], torch.from_numpy(gold['video_deformed'])) < 1e-4


def test_synthetic_inputs_are_reproducible():
    gold = helpers.load_golden('golden_tiny')
    assert torch.equal(helpers.smooth_frames(2, 1, 32, 5), torch.from_numpy(gold['source']))


@pytest.mark.skipif(not ref_shim.available(), reason='/root/reference only exists in the build container')
@pytest.mark.parametrize('name,res', [('moving-gif', 64), ('vox-full', 128), ('bair', 64), ('taichi', 64), ('shapes', 64),
                                      ('nemo', 64), ('vox', 128), ('actions', 64)])
def test_oracle_matches_live_reference(name, res):
    """Forward parity with identical keypoints fed to both (the chain is ill-conditioned w.r.t. 1e-7 kp noise)."""
    cfg = helpers.load_config(name)
    torch.manual_seed(0)
    rg, rd, rk = ref_shim.build_from_config(cfg)
    og, od, ok = mo.build_from_config(cfg)
    helpers.perturb_flow_head(rg)
    og.load_state_dict(rg.state_dict()); od.load_state_dict(rd.state_dict()); ok.load_state_dict(rk.state_dict())
    x = {'source': helpers.smooth_frames(1, 1, res, 5), 'video': helpers.smooth_frames(1, 2, res, 6)}
    for m in (rg, rd, rk, og, od, ok):
        m.eval()
    with torch.no_grad():
        a, b = rk(x['video']), ok(x['video'])
        assert helpers.max_abs(a['mean'], b['mean']) < 1e-5 and helpers.max_abs(a['var'], b['var']) < 1e-5
        ks = {k: v[:, :1] for k, v in a.items()}
        ra, oa = rg(x['source'], kp_driving=a, kp_source=ks), og(x['source'], kp_driving=a, kp_source=ks)
        assert ra['video_prediction'].shape == oa['video_prediction'].shape
        assert helpers.max_abs(ra['video_prediction'], oa['video_prediction']) < 1e-4
        assert helpers.max_abs(ra['video_deformed'], oa['video_deformed']) < 1e-4
        kd1 = {k: v[:, :1] for k, v in a.items()}
        rm, om = rd(x['video'][:, :, :1], kd1, ks), od(x['video'][:, :, :1], kd1, ks)
        assert max(helpers.max_abs(p, q) for p, q in zip(rm, om)) < 1e-4
