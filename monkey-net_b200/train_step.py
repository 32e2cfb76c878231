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


def train_iteration(generator_full_par, discriminator_full_par, optimizers, train_params, x):
    """One loop body of train.py:110-136.  Returns (generator loss tensors, discriminator loss tensors)."""
    opt_g, opt_d, opt_kp = optimizers
    out = generator_full_par(x)
    loss_values = [val.mean() for val in out[:-2]]
    generated, kp_joined = out[-2], out[-1]
    loss = sum(loss_values)
    loss.backward(retain_graph=not train_params['detach_kp_discriminator'])
    opt_g.step(); opt_g.zero_grad(); opt_d.zero_grad()
    if train_params['detach_kp_discriminator']:
        opt_kp.step(); opt_kp.zero_grad()
    g_vals = loss_values
    loss_values = [val.mean() for val in discriminator_full_par(x, kp_joined, generated)]
    loss = sum(loss_values)
    loss.backward()
    opt_d.step(); opt_d.zero_grad()
    if not train_params['detach_kp_discriminator']:
        opt_kp.step(); opt_kp.zero_grad()
    return g_vals, loss_values
