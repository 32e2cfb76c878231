"""Host-side motion-transfer composition, mirroring the reference's transfer.py so callers that cannot import the
reference driver (it needs imageio / skimage / matplotlib) can run the exact inference path:

  normalize_kp    == transfer.py:31-62  (move_location, movement_mult, clip_mean, adapt_variance)
  transfer_one    == transfer.py:65-79  (KP detector on every driving frame, generator on every frame)
  reconstruct     == reconstruction.py:12-25,57-62 (frame 0 is the appearance, every frame drives)

Differences, both result-preserving in eval mode (where transfer runs, transfer.py:104-105):
  * `batched=True` (default) feeds all d driving frames through ONE keypoint-detector call (D = d) and ONE generator
    call: the appearance encoder of the source runs once instead of d times and every kernel sees B*d frames
    (SURVEY 8(f) rank 3).  Eval-mode batch norm uses running statistics, so frames stay independent and the output
    equals the reference's per-frame loop (`batched=False` is that loop, verbatim).
  * `torch.gesv` (removed from torch) is replaced by the closed-form 2x2 inverse, the numpy eig of the symmetrised
    variance by `torch.linalg.eigh` of the same matrix.
`GraphedTransfer` replays the whole call as one CUDA graph for fixed shapes.
"""
import numpy as np
import torch

from modules.util import matrix_inverse


def _cat_dict(dicts, dim):
    return {k: torch.cat([v[k] for v in dicts], dim=dim) for k in dicts[0]}


def make_symetric_matrix(m):
    """transfer.py:17-28: symmetrise, floor non-positive eigenvalues at 1e-6, recompose."""
    c = (m + m.transpose(-1, -2)) / 2
    d, u = torch.linalg.eigh(c)
    d = torch.where(d <= 0, torch.full_like(d, 1e-6), d)
    return (u * d.unsqueeze(-2)) @ u.transpose(-1, -2)


def normalize_kp(kp_video, kp_appearance, movement_mult=False, move_location=False, adapt_variance=False,
                 clip_mean=False):
    if movement_mult:
        from scipy.spatial import ConvexHull
        appearance_area = ConvexHull(kp_appearance['mean'][0, 0].detach().cpu().numpy()).volume
        video_area = ConvexHull(kp_video['mean'][0, 0].detach().cpu().numpy()).volume
        movement_mult = float(np.sqrt(appearance_area) / np.sqrt(video_area))
    else:
        movement_mult = 1
    kp_video = {k: v for k, v in kp_video.items()}
    if move_location:
        kp_video_diff = (kp_video['mean'] - kp_video['mean'][:, 0:1]) * movement_mult
        kp_video['mean'] = kp_video_diff + kp_appearance['mean']
    if clip_mean:
        kp_video['mean'] = kp_video['mean'].clamp(-1.0, 1.0)
    if ('var' in kp_video) and adapt_variance:
        kp_var = torch.matmul(kp_video['var'], matrix_inverse(kp_video['var'][:, 0:1], eps=0))
        kp_var = torch.matmul(kp_var, kp_appearance['var'])
        kp_video['var'] = make_symetric_matrix(kp_var)
    return kp_video


def transfer_one(generator, kp_detector, source_image, driving_video, transfer_params, batched=True):
    """source_image (B,C,1,H,W), driving_video (B,C,d,H,W) -> dict with 'video_prediction' / 'video_deformed'
    (B,C,d,H,W) and the keypoints ('kp_driving', 'kp_source', 'kp_norm'), exactly the reference's dictionary."""
    d = driving_video.shape[2]
    norm = transfer_params.get('normalization_params', {}) if 'normalization_params' in transfer_params \
        else transfer_params
    if batched:
        kp_driving = kp_detector(driving_video)
    else:
        kp_driving = _cat_dict([kp_detector(driving_video[:, :, i:(i + 1)]) for i in range(d)], dim=1)
    kp_source = kp_detector(source_image)
    kp_driving_norm = normalize_kp(kp_driving, kp_source, **norm)
    if batched:
        out = dict(generator(source_image=source_image, kp_driving=kp_driving_norm, kp_source=kp_source))
    else:
        kp_video_list = [{k: v[:, i:(i + 1)] for k, v in kp_driving_norm.items()} for i in range(d)]
        out = _cat_dict([generator(source_image=source_image, kp_driving=kp, kp_source=kp_source)
                         for kp in kp_video_list], dim=2)
    out['kp_driving'] = kp_driving
    out['kp_source'] = kp_source
    out['kp_norm'] = kp_driving_norm
    return out


def reconstruct(generator, kp_detector, video, batched=True):
    """reconstruction.py:12-25,57-62: frame 0 of `video` (B,C,d,H,W) is the appearance, every frame is a driving
    frame.  Returns the reference's dictionary: 'video_prediction' / 'video_deformed' (B,C,d,H,W), 'kp_driving',
    'kp_source'.  `batched` as in transfer_one (eval mode: identical results, one KP pass + one generator pass)."""
    d = video.shape[2]
    source = video[:, :, :1]
    kp_source = kp_detector(source)
    if batched:
        kp_video = kp_detector(video)
        out = dict(generator(source, kp_driving=kp_video, kp_source=kp_source))
    else:
        kp_video = _cat_dict([kp_detector(video[:, :, i:(i + 1)]) for i in range(d)], dim=1)
        parts = [generator(source, kp_driving={k: v[:, i:(i + 1)] for k, v in kp_video.items()}, kp_source=kp_source)
                 for i in range(d)]
        out = {k: torch.cat([p[k] for p in parts], dim=2) for k in ('video_prediction', 'video_deformed')}
    out['kp_driving'] = kp_video
    out['kp_source'] = kp_source
    return out


class GraphedTransfer:
    """transfer_one for fixed shapes as ONE CUDA graph launch: `run(source, driving)` copies the inputs into static
    buffers (H2D when they live in pinned host memory) and replays; returns the static 'video_prediction' tensor
    (B,C,d,H,W).  The modules must be in eval mode (running statistics are read, never written)."""

    def __init__(self, generator, kp_detector, transfer_params, use_graph=True, warmup=2):
        self.generator, self.kp_detector, self.transfer_params = generator, kp_detector, transfer_params
        self.device = next(generator.parameters()).device
        self.use_graph, self.warmup = bool(use_graph), warmup
        self.graph = None
        self.static_in = self.static_out = None
        self.kernels_per_call = 0
        norm = transfer_params.get('normalization_params', transfer_params)
        if use_graph and norm.get('movement_mult'):
            raise ValueError('movement_mult runs scipy ConvexHull on the host and cannot be captured in a CUDA graph')

    def _call(self, source, driving):
        with torch.no_grad():
            return transfer_one(self.generator, self.kp_detector, source, driving, self.transfer_params)

    def _capture(self, source, driving):
        assert not self.generator.training and not self.kp_detector.training, 'transfer runs in eval mode'
        self.static_in = (torch.empty(source.shape, dtype=source.dtype, device=self.device),
                          torch.empty(driving.shape, dtype=driving.dtype, device=self.device))
        self.static_in[0].copy_(source)
        self.static_in[1].copy_(driving)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._call(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import lib
        graph = torch.cuda.CUDAGraph()
        n0 = lib.launches()
        with torch.cuda.graph(graph):
            self.static_out = self._call(*self.static_in)
        self.kernels_per_call = lib.launches() - n0
        self.graph = graph
        # The captured kernels read the CACHED weight packs / folded eval-BN vectors (ops._INFER_CACHE), not the live
        # parameters: keep those tensors alive for the life of the graph (a later eager call may replace the cache
        # entries, and a freed pack would be recycled by the allocator under the graph's feet) and remember the
        # parameter state they were built from.
        from . import ops
        self._pinned = list(ops._INFER_CACHE.values())
        self._state = self._param_state()

    def _param_state(self):
        from . import ops
        vers = 0
        for m in (self.generator, self.kp_detector):
            for t in list(m.parameters()) + list(m.buffers()):
                vers += t._version
        return (ops._PARAM_EPOCH[0], vers)

    def run(self, source, driving):
        if not self.use_graph:
            return self._call(source.to(self.device, non_blocking=True), driving.to(self.device, non_blocking=True))
        if self.graph is not None and self._param_state() != self._state:
            self.graph = None  # load_state_dict / a training step / an optimiser update since capture: stale packs
        if self.graph is None:
            self._capture(source, driving)
        self.static_in[0].copy_(source, non_blocking=True)
        self.static_in[1].copy_(driving, non_blocking=True)
        self.graph.replay()
        return self.static_out
