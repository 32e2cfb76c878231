"""Host-side training-step composition, mirroring the reference's train.py so bench.py / smoke() can run the exact
iteration body without importing the reference's driver (which needs logger/skimage/imageio):

  GeneratorFullModel / DiscriminatorFullModel  == train.py:24-75
  train_iteration                              == train.py:110-136 (logging excluded)

The reference's own, unmodified train.py runs against `modules/` + `sync_batchnorm/` the same way.
"""
import torch

from modules.losses import generator_loss, discriminator_loss


def split_kp(kp_joined, detach=False):
    f = (lambda t: t.detach()) if detach else (lambda t: t)
    return {'kp_driving': {k: f(v[:, 1:]) for k, v in kp_joined.items()},
            'kp_source': {k: f(v[:, :1]) for k, v in kp_joined.items()}}


class GeneratorFullModel(torch.nn.Module):
    def __init__(self, kp_extractor, generator, discriminator, train_params):
        super(GeneratorFullModel, self).__init__()
        self.kp_extractor, self.generator, self.discriminator = kp_extractor, generator, discriminator
        self.train_params = train_params

    def forward(self, x):
        kp_joined = self.kp_extractor(torch.cat([x['source'], x['video']], dim=2))
        generated = self.generator(x['source'], **split_kp(kp_joined, self.train_params['detach_kp_generator']))
        kp_dict = split_kp(kp_joined, False)
        maps_gen = self.discriminator(generated['video_prediction'], **kp_dict)
        maps_real = self.discriminator(x['video'], **kp_dict)
        generated.update(kp_dict)
        losses = generator_loss(discriminator_maps_generated=maps_gen, discriminator_maps_real=maps_real,
                                video_deformed=generated['video_deformed'],
                                loss_weights=self.train_params['loss_weights'])
        return tuple(losses) + (generated, kp_joined)


class DiscriminatorFullModel(torch.nn.Module):
    def __init__(self, kp_extractor, generator, discriminator, train_params):
        super(DiscriminatorFullModel, self).__init__()
        self.kp_extractor, self.generator, self.discriminator = kp_extractor, generator, discriminator
        self.train_params = train_params

    def forward(self, x, kp_joined, generated):
        kp_dict = split_kp(kp_joined, self.train_params['detach_kp_discriminator'])
        maps_gen = self.discriminator(generated['video_prediction'].detach(), **kp_dict)
        maps_real = self.discriminator(x['video'], **kp_dict)
        return discriminator_loss(discriminator_maps_generated=maps_gen, discriminator_maps_real=maps_real,
                                  loss_weights=self.train_params['loss_weights'])


def make_optimizers(generator, discriminator, kp_detector, lr):
    mk = lambda m: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.5, 0.999))
    return mk(generator), mk(discriminator), mk(kp_detector)


def train_iteration(generator_full_par, discriminator_full_par, optimizers, train_params, x, grad_sync=None):
    """One loop body of train.py:110-136.  Returns (generator loss tensors, discriminator loss tensors).
    `optimizers` are torch.optim.Adam objects (the reference's code verbatim; with `grad_sync` = (sync_g, sync_d,
    sync_kp) FlatGradSync objects for data-parallel runs: one all-reduce per optimiser step, gradients zeroed in place
    like torch 0.4.1 did) or monkey_net_b200.optim.FlatAdam objects (one fused launch per step that also zeroes its
    gradients and owns the flat all-reduce)."""
    from .optim import FlatAdam
    opt_g, opt_d, opt_kp = optimizers
    if isinstance(opt_g, FlatAdam):
        return _train_iteration_flat(generator_full_par, discriminator_full_par, optimizers, train_params, x)
    keep = grad_sync is not None
    zero = (lambda o: o.zero_grad(set_to_none=False)) if keep else (lambda o: o.zero_grad())
    sync = (lambda i: grad_sync[i].sync()) if keep else (lambda i: None)
    out = generator_full_par(x)
    loss_values = [val.mean() for val in out[:-2]]
    generated, kp_joined = out[-2], out[-1]
    loss = sum(loss_values)
    loss.backward(retain_graph=not train_params['detach_kp_discriminator'])
    sync(0); opt_g.step(); zero(opt_g); zero(opt_d)
    if train_params['detach_kp_discriminator']:
        sync(2); opt_kp.step(); zero(opt_kp)
    g_vals = loss_values
    loss_values = [val.mean() for val in discriminator_full_par(x, kp_joined, generated)]
    loss = sum(loss_values)
    loss.backward()
    sync(1); opt_d.step(); zero(opt_d)
    if not train_params['detach_kp_discriminator']:
        sync(2); opt_kp.step(); zero(opt_kp)
    return g_vals, loss_values


_COMM_STREAM = {}


def _comm_stream(device):
    if device not in _COMM_STREAM:
        _COMM_STREAM[device] = torch.cuda.Stream(device=device)
    return _COMM_STREAM[device]


def _train_iteration_flat(generator_full_par, discriminator_full_par, optimizers, train_params, x):
    """train.py:110-136 with FlatAdam: `step()` = all-reduce (N > 1) + update + zero_grad of that group in one launch.

    Data parallel (N > 1): the discriminator step reads the generated frames and keypoints DETACHED
    (train.py:71, detach_kp_discriminator) - it depends on neither the generator's nor the keypoint detector's UPDATED
    parameters.  Their flat gradient all-reduces (NVLink, ~170 MB at taichi) and Adam launches therefore run on a side
    stream concurrently with the whole discriminator step and are joined at the end of the iteration - the collective
    is off the critical path (fork / join is captured in the CUDA graph as stream dependencies)."""
    from . import dist as mkdist
    from . import ops as mkops
    opt_g, opt_d, opt_kp = optimizers
    out = generator_full_par(x)
    loss_values = [val.mean() for val in out[:-2]]
    generated, kp_joined = out[-2], out[-1]
    loss = sum(loss_values)
    with mkops.direct_param_grads():   # conv weight gradients accumulate straight into FlatAdam's gradient buffers
        loss.backward(retain_graph=not train_params['detach_kp_discriminator'])
    overlap = mkdist.world() > 1 and train_params['detach_kp_discriminator'] and loss.is_cuda
    side = None
    if overlap:
        main = torch.cuda.current_stream()
        side = _comm_stream(loss.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            opt_g.sync_gradients(); opt_g.step()
            opt_kp.sync_gradients(); opt_kp.step()
    else:
        opt_g.sync_gradients(); opt_g.step()      # optimizer_generator.step(); .zero_grad()
    opt_d.zero_grad()                         # optimizer_discriminator.zero_grad()
    if train_params['detach_kp_discriminator'] and not overlap:
        opt_kp.sync_gradients(); opt_kp.step()
    g_vals = loss_values
    loss_values = [val.mean() for val in discriminator_full_par(x, kp_joined, generated)]
    loss = sum(loss_values)
    with mkops.direct_param_grads():
        loss.backward()
    opt_d.sync_gradients(); opt_d.step()
    if not train_params['detach_kp_discriminator']:
        opt_kp.sync_gradients(); opt_kp.step()
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)   # join: the next iteration reads the updated G / KP parameters
    return g_vals, loss_values


class GraphedTrainer:
    """The training iteration as ONE CUDA graph launch (B200 design rule: streams and graphs, not a tracing compiler).

    At 64x64 the step is ~850 small kernels and the eager loop is bound by Python / launch latency, not by the GPU.
    `GraphedTrainer` runs `train_iteration` eagerly a few times on a side stream (with parameters, BN statistics and
    optimiser state snapshotted and restored afterwards, so warm-up does not advance training), captures one
    iteration - forward, backward, the three Adam steps, and the NCCL collectives when a process group exists - into a
    `torch.cuda.CUDAGraph`, and from then on `step(x)` is: copy `x` into the static input buffers (H2D when `x` lives
    in pinned host memory) + one graph replay.  Same arithmetic, same kernels, same update order as train.py:110-136.
    """

    def __init__(self, kp_detector, generator, discriminator, train_params, use_graph=True, warmup=3,
                 fused_adam=True):
        from sync_batchnorm import DataParallelWithCallback
        self.modules = (kp_detector, generator, discriminator)
        self.train_params = train_params
        self.device = next(generator.parameters()).device
        ids = [self.device.index] if self.device.index is not None else None
        self.g_full = DataParallelWithCallback(GeneratorFullModel(kp_detector, generator, discriminator, train_params),
                                               device_ids=ids)
        self.d_full = DataParallelWithCallback(DiscriminatorFullModel(kp_detector, generator, discriminator,
                                                                      train_params), device_ids=ids)
        self.grad_sync = None
        from . import dist as mkdist
        if mkdist.world() > 1:
            # data parallel: ONE all-reduce per optimiser step over a flat gradient buffer instead of the wrapper's
            # per-parameter hooks (about 200 latency-bound collectives per iteration)
            import sync_batchnorm.replicate as rep
            rep.HOOKS_ENABLED = False
        self.fused_adam = bool(fused_adam)
        if self.fused_adam:
            from .optim import FlatAdam
            mk = lambda m: FlatAdam(m.parameters(), lr=train_params['lr'], betas=(0.5, 0.999))
            self.optimizers = (mk(generator), mk(discriminator), mk(kp_detector))
        else:
            mk = lambda m: torch.optim.Adam(m.parameters(), lr=train_params['lr'], betas=(0.5, 0.999),
                                            capturable=bool(use_graph))
            self.optimizers = (mk(generator), mk(discriminator), mk(kp_detector))
            if mkdist.world() > 1:
                self.grad_sync = (mkdist.FlatGradSync(generator.parameters()),
                                  mkdist.FlatGradSync(discriminator.parameters()),
                                  mkdist.FlatGradSync(kp_detector.parameters()))
        # MultiStepLR of train.py:92-97 (one scheduler per optimiser, stepped once per epoch, train.py:146-148)
        self.schedulers = None
        ms = train_params.get('epoch_milestones') if hasattr(train_params, 'get') else None
        if ms is not None:
            if self.fused_adam:
                from .optim import MultiStepLR
            else:
                from torch.optim.lr_scheduler import MultiStepLR
            self.schedulers = [MultiStepLR(o, ms, gamma=0.1) for o in self.optimizers]
        self.use_graph, self.warmup = bool(use_graph), warmup
        self.graph = None
        self.static_in = self.static_out = None
        self.kernels_per_step = 0

    def _iteration(self, x):
        g_vals, d_vals = train_iteration(self.g_full, self.d_full, self.optimizers, self.train_params, x,
                                         self.grad_sync)
        return torch.stack([v.detach() for v in g_vals + d_vals])

    def _snapshot(self):
        import copy
        return ([copy.deepcopy(m.state_dict()) for m in self.modules],
                [copy.deepcopy(o.state_dict()) for o in self.optimizers])

    def _restore(self, snap):
        for m, sd in zip(self.modules, snap[0]):
            m.load_state_dict(sd)
        for o, sd in zip(self.optimizers, snap[1]):
            if self.fused_adam:
                o.load_state_dict(sd)
                continue
            if len(sd['state']) == 0:
                # fresh optimiser: keep the (now allocated) moment / step tensors - the graph must capture only the
                # update, not their lazy zero-initialisation - and reset their values in place
                for st in o.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            else:
                o.load_state_dict(sd)

    def _capture(self, x):
        self.static_in = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in x.items()}
        for k, v in x.items():
            self.static_in[k].copy_(v)
        snap = self._snapshot()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._iteration(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        from . import lib
        graph = torch.cuda.CUDAGraph()
        n0 = lib.launches()
        with torch.cuda.graph(graph):
            self.static_out = self._iteration(self.static_in)
        self.kernels_per_step = lib.launches() - n0  # C-ABI kernel launches recorded into the graph
        self.graph = graph

    def epoch_end(self):
        """scheduler_*.step() of train.py:146-148.  With FlatAdam the rate is a device scalar, so the captured graph
        follows the schedule; torch.optim.Adam(capturable=True) keeps lr as a tensor only if it was created as one -
        that path re-captures."""
        if self.schedulers is None:
            return
        for s in self.schedulers:
            s.step()
        if not self.fused_adam and self.graph is not None:
            self.graph = None  # host-float lr baked into the captured foreach kernels: capture again

    def step(self, x):
        """x = {'source': (B,3,1,H,W), 'video': (B,3,1,H,W)} on this device or in (pinned) host memory.  Returns the
        loss values of the iteration as one device tensor [generator terms..., discriminator term]."""
        if not self.use_graph:
            return self._iteration(x)
        if self.graph is None:
            self._capture(x)
        for k, v in x.items():
            self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        from . import ops
        ops.note_parameters_changed()  # the replay updated parameters and running statistics without any Python
        return self.static_out
