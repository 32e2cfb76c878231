"""Data edge: the stacked-frame ingestion kernel vs the reference's read_video arithmetic
(frames_dataset.py:14-29, restated in numpy in monkey_net_b200/data.py and pinned to the reference formula on CPU in
tests/test_data_host.py).  Bit-exact: uint8 / 255 in float32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('cs', [0, 1, 2, 3, 4])
@pytest.mark.parametrize('h,w,T', [(64, 64, 32), (21, 21, 5)])   # square, as in every shipped config
def test_stacked_u8_to_nhwc_bit_exact(cs, h, w, T):
    from monkey_net_b200 import data
    rng = np.random.default_rng(cs * 100 + h)
    shape = (h, T * w) if cs == 0 else (h, T * w, cs)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    video, nhwc = data.stacked_to_device(img, (h, w, 3))
    want = data.read_video_reference_semantics(img if cs != 2 else img, (h, w, 3))          # (T, h, w, 3)
    got = video[0].permute(1, 2, 3, 0).cpu().numpy()                                         # (T, h, w, 3)
    assert got.shape == want.shape == (T, h, w, 3)
    assert np.array_equal(got, want)
    assert float(nhwc[..., 3].abs().max()) == 0.0


def test_ingested_video_feeds_the_modules():
    """the (1,3,T,H,W) view is a legal module input (NHWC-backed, non-contiguous)"""
    import helpers
    import test_gpu_2_modules as t2
    from monkey_net_b200 import data
    cfg = helpers.tiny_config()
    gen, disc, kp = t2.build_product(cfg)
    kp.cuda().eval()
    img = (helpers.smooth_frames(1, 4, 32, 3)[0].permute(2, 1, 3, 0).reshape(32, 4 * 32, 3) * 255).round().to(torch.uint8)
    video, _ = data.stacked_to_device(img.numpy(), (32, 32, 3))
    with torch.no_grad():
        a = kp(video)
        b = kp(video.contiguous())
    assert a['mean'].shape == (1, 4, cfg['model_params']['common_params']['num_kp'], 2)
    assert float((a['mean'] - b['mean']).abs().max()) < 1e-6   # split-K atomics: equal up to summation order
