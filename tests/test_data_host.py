"""Host side of the data edge: the numpy restatement of frames_dataset.read_video's arithmetic (used as the oracle of the
ingestion kernel) against an independent derivation, on a PNG written and decoded with PIL - stacked-frame layout, gray
and RGBA handling, uint8/255 in float32.  (The reference function itself needs scikit-image, absent here.)"""
import io

import numpy as np


def test_restatement_matches_the_reference_formula_on_a_decoded_png():
    from PIL import Image
    from monkey_net_b200 import data
    rng = np.random.default_rng(0)
    T, h, w = 6, 16, 16     # square frames: every shipped config (the reference's reshape only works for them)
    frames = rng.integers(0, 256, size=(T, h, w, 4), dtype=np.uint8)
    stacked = np.concatenate(list(frames), axis=1)                     # (h, T*w, 4): frames side by side
    buf = io.BytesIO()
    Image.fromarray(stacked, 'RGBA').save(buf, format='PNG')
    decoded = np.array(Image.open(io.BytesIO(buf.getvalue())))
    assert np.array_equal(decoded, stacked)
    out = data.read_video_reference_semantics(decoded, (h, w, 3))
    assert out.shape == (T, h, w, 3) and out.dtype == np.float32
    want = frames[..., :3].astype(np.float32) / np.float32(255)
    assert np.array_equal(out, want)
    gray = data.read_video_reference_semantics(stacked[..., 0], (h, w, 3))
    assert np.array_equal(gray, np.repeat(want[..., :1], 3, axis=3))
