"""Losses - drop-in for the reference's `modules/losses.py`: feature-matching L1 over the discriminator maps + LSGAN
terms, each a per-sample mean of shape (B,) (train.py:114 takes `.mean()` of every entry).  Each term is one
stride-aware reduction kernel (forward) and one elementwise kernel (backward)."""
import torch

from monkey_net_b200 import ops


def mean_batch(val):
    return val.reshape(val.shape[0], -1).mean(-1)


def reconstruction_loss(prediction, target, weight):
    if weight == 0:
        return 0
    return ops.loss_mean('l1', prediction, target, weight)


def generator_gan_loss(discriminator_maps_generated, weight):
    return ops.loss_mean('gen_gan', discriminator_maps_generated[-1], None, weight)


def discriminator_gan_loss(discriminator_maps_generated, discriminator_maps_real, weight):
    return ops.loss_mean('disc_gan', discriminator_maps_real[-1], discriminator_maps_generated[-1], weight)


def generator_loss_names(loss_weights):
    loss_names = []
    if loss_weights['reconstruction_deformed'] != 0:
        loss_names.append("rec_def")
    if loss_weights['reconstruction'] is not None:
        for i, _ in enumerate(loss_weights['reconstruction']):
            if loss_weights['reconstruction'][i] == 0:
                continue
            loss_names.append("layer-%s_rec" % i)
    loss_names.append("gen_gan")
    return loss_names


def discriminator_loss_names():
    return ['disc_gan']


def generator_loss(discriminator_maps_generated, discriminator_maps_real, video_deformed, loss_weights):
    loss_values = []
    if loss_weights['reconstruction_deformed'] != 0:
        loss_values.append(reconstruction_loss(discriminator_maps_real[0], video_deformed,
                                               loss_weights['reconstruction_deformed']))
    if loss_weights['reconstruction'] != 0:
        for i, (real, gen) in enumerate(zip(discriminator_maps_real[:-1], discriminator_maps_generated[:-1])):
            if loss_weights['reconstruction'][i] == 0:
                continue
            loss_values.append(reconstruction_loss(gen, real, weight=loss_weights['reconstruction'][i]))
    loss_values.append(generator_gan_loss(discriminator_maps_generated, weight=loss_weights['generator_gan']))
    return loss_values


def discriminator_loss(discriminator_maps_generated, discriminator_maps_real, loss_weights):
    return [discriminator_gan_loss(discriminator_maps_generated, discriminator_maps_real,
                                   loss_weights['discriminator_gan'])]
