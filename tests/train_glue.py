"""The reference's train.py FullModel compositions (train.py:24-75), restated for tests so they can drive either the
oracle or the product modules (train.py itself imports logger/skimage, which this image lacks)."""
import torch


def split_kp(kp_joined, detach=False):
    f = (lambda t: t.detach()) if detach else (lambda t: t)
    return {'kp_driving': {k: f(v[:, 1:]) for k, v in kp_joined.items()},
            'kp_source': {k: f(v[:, :1]) for k, v in kp_joined.items()}}


def generator_full(kp_extractor, generator, discriminator, train_params, x):
    from modules.losses import generator_loss
    kp_joined = kp_extractor(torch.cat([x['source'], x['video']], dim=2))
    generated = generator(x['source'], **split_kp(kp_joined, train_params['detach_kp_generator']))
    kp_dict = split_kp(kp_joined, False)
    maps_gen = discriminator(generated['video_prediction'], **kp_dict)
    maps_real = discriminator(x['video'], **kp_dict)
    generated.update(kp_dict)
    losses = generator_loss(discriminator_maps_generated=maps_gen, discriminator_maps_real=maps_real,
                            video_deformed=generated['video_deformed'], loss_weights=train_params['loss_weights'])
    return tuple(losses) + (generated, kp_joined)


def discriminator_full(kp_extractor, generator, discriminator, train_params, x, kp_joined, generated):
    from modules.losses import discriminator_loss
    kp_dict = split_kp(kp_joined, train_params['detach_kp_discriminator'])
    maps_gen = discriminator(generated['video_prediction'].detach(), **kp_dict)
    maps_real = discriminator(x['video'], **kp_dict)
    return discriminator_loss(discriminator_maps_generated=maps_gen, discriminator_maps_real=maps_real,
                              loss_weights=train_params['loss_weights'])
