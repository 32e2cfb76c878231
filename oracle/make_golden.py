"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (/root/reference) on CPU.

Run in the build container only (`python oracle/make_golden.py`); the fixtures travel to the GPU box, the
reference does not.  Each fixture holds the inputs' recipe, the reference outputs of one training-mode step
(keypoints, generator outputs, losses, per-parameter gradient norms) and one eval-mode forward.
  golden_tiny.npz    small architecture, FULL reference weights stored in the file
  golden_shapes.npz  config/shapes.yaml, weights = default init under torch.manual_seed(0) built in run.py order
                     (generator, discriminator, kp_detector) - the product reproduces them bit-exactly and the
                     fixture stores a checksum to prove it.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import ref_shim  # noqa: E402
import helpers  # noqa: E402


def checksum(sd):
    tot = 0.0
    for i, (k, v) in enumerate(sorted(sd.items())):
        tot += float(v.double().abs().sum()) * (1 + (i % 7))
    return tot


def run_reference(cfg, res, batch, store_weights):
    ref = ref_shim.load()
    torch.manual_seed(0)
    gen, disc, kp = ref_shim.build_from_config(cfg)
    helpers.perturb_flow_head(gen)
    out = {}
    if store_weights:
        for tag, m in (('G', gen), ('D', disc), ('K', kp)):
            for k, v in m.state_dict().items():
                out['w/%s/%s' % (tag, k)] = v.numpy().copy()
    out['checksum'] = np.array([checksum(gen.state_dict()), checksum(disc.state_dict()), checksum(kp.state_dict())])
    x = {'source': helpers.smooth_frames(batch, 1, res, 5), 'video': helpers.smooth_frames(batch, 1, res, 6)}
    out['source'], out['video'] = x['source'].numpy(), x['video'].numpy()
    tp = cfg['train_params']
    for m in (gen, disc, kp):
        m.train()
    kp_joined = kp(torch.cat([x['source'], x['video']], dim=2))
    out['kp_mean'], out['kp_var'] = kp_joined['mean'].detach().numpy(), kp_joined['var'].detach().numpy()
    split = lambda kj: ({k: v[:, 1:] for k, v in kj.items()}, {k: v[:, :1] for k, v in kj.items()})
    kd, ks = split(kp_joined)
    generated = gen(x['source'], kp_driving=kd, kp_source=ks)
    out['video_prediction'] = generated['video_prediction'].detach().numpy()
    out['video_deformed'] = generated['video_deformed'].detach().numpy()
    maps_gen = disc(generated['video_prediction'], kp_driving=kd, kp_source=ks)
    maps_real = disc(x['video'], kp_driving=kd, kp_source=ks)
    losses = ref['losses'].generator_loss(maps_gen, maps_real, generated['video_deformed'], tp['loss_weights'])
    out['g_losses'] = np.stack([l.detach().numpy() for l in losses])
    out['disc_score_gen'] = maps_gen[-1].detach().numpy()
    sum(l.mean() for l in losses).backward()
    for tag, m in (('G', gen), ('D', disc), ('K', kp)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                out['gnorm/%s/%s' % (tag, k)] = np.array(float(p.grad.norm()))
    kdd, ksd = split({k: v.detach() for k, v in kp_joined.items()})
    d_loss = ref['losses'].discriminator_loss(disc(generated['video_prediction'].detach(), kp_driving=kdd, kp_source=ksd),
                                              disc(x['video'], kp_driving=kdd, kp_source=ksd), tp['loss_weights'])
    out['d_losses'] = np.stack([l.detach().numpy() for l in d_loss])
    # eval-mode forward with the running statistics the training forward just produced
    for m in (gen, disc, kp):
        m.eval()
    with torch.no_grad():
        kpe = kp(torch.cat([x['source'], x['video']], dim=2))
        kd, ks = split(kpe)
        ge = gen(x['source'], kp_driving=kd, kp_source=ks)
    out['eval_kp_mean'] = kpe['mean'].numpy()
    out['eval_video_prediction'] = ge['video_prediction'].numpy()
    return out


def read_strip_png(path, image_shape=(64, 64, 3)):
    """frames_dataset.py:14-29 (read_video, .png branch) restated with PIL: a W*T x H strip -> (T, H, W, 3) float32."""
    from PIL import Image
    image = np.array(Image.open(path))
    if image.ndim == 2:
        image = np.stack([image] * 3, -1)
    if image.shape[2] == 4:
        image = image[..., :3]
    video = np.moveaxis(image, 1, 0).reshape((-1,) + image_shape)
    return np.moveaxis(video, 1, 2)  # uint8; img_as_float32 == / 255


def run_reconstruction(cfg, png):
    """BASELINE.json configs[0] (PR1): config/shapes.yaml as shipped, eval mode, seed-0 default-init weights (+ the
    seeded flow-head perturbation so the warp is not the identity), reconstruction of one bundled data/shapes test
    video (32 frames 64x64, B=1) by the reference's OWN `generate` (reconstruction.py:12-25, body extracted from the
    unmodified file) around the reference's own modules, keypoints as in reconstruction.py:57-62."""
    import ast
    torch.manual_seed(0)
    gen, disc, kp = ref_shim.build_from_config(cfg)
    helpers.perturb_flow_head(gen)
    src = open(os.path.join(ref_shim.REF_ROOT, 'reconstruction.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'generate']
    ns = {'torch': torch}
    exec(compile(ast.Module(body=fn, type_ignores=[]), 'reference/reconstruction.py', 'exec'), ns)
    frames = read_strip_png(png)
    video = torch.from_numpy(frames.astype(np.float32) / 255.0).permute(3, 0, 1, 2)[None].contiguous()  # (1,3,T,H,W)
    for m in (gen, kp):
        m.eval()
    cat_dict = lambda l, dim: {k: torch.cat([v[k] for v in l], dim=dim) for k in l[0]}
    with torch.no_grad():
        kp_appearance = kp(video[:, :, :1])
        kp_video = cat_dict([kp(video[:, :, i:(i + 1)]) for i in range(video.shape[2])], dim=1)
        out = ns['generate'](gen, appearance_image=video[:, :, :1], kp_appearance=kp_appearance, kp_video=kp_video)
    keep = [0, 7, 16, 31]
    return {'frames_u8': frames, 'checksum': np.array([checksum(gen.state_dict()), checksum(kp.state_dict())]),
            'kp_mean': kp_video['mean'].numpy(), 'kp_var': kp_video['var'].numpy(), 'keep': np.array(keep),
            'video_prediction': out['video_prediction'][:, :, keep].numpy(),
            'video_deformed': out['video_deformed'][:, :, keep].numpy()}


def main():
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    recon = run_reconstruction(helpers.load_config('shapes'),
                               os.path.join(ref_shim.REF_ROOT, 'data', 'shapes', 'test', '00000001.png'))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'golden_recon_shapes.npz'), **recon)
    print('golden_recon_shapes.npz', os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'golden_recon_shapes.npz')) // 1024, 'KiB')
    if '--recon-only' in sys.argv:
        return
    tiny = run_reference(helpers.tiny_config(), 32, 2, True)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'golden_tiny.npz'), **tiny)
    shapes = run_reference(helpers.load_config('shapes'), 64, 2, False)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'golden_shapes.npz'), **shapes)
    for n in ('golden_tiny.npz', 'golden_shapes.npz'):
        print(n, os.path.getsize(os.path.join(ROOT, 'tests', 'golden', n)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
