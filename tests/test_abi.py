"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/monkey_b200.h
declares, the ctypes binding is derived from that header, and the Python module API matches the reference's."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import helpers

ROOT = helpers.ROOT


@pytest.fixture(scope='module')
def built_lib():
    import __graft_entry__ as g
    g.build()
    from monkey_net_b200 import lib
    return lib


def test_library_exports_every_declared_symbol(built_lib):
    text = open(os.path.join(ROOT, 'include', 'monkey_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    declared = set(re.findall(r'\b(mk_\w+)\s*\(', text))
    assert len(declared) >= 30
    cdll = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert hasattr(cdll, name), 'symbol %s declared in include/monkey_b200.h but not exported' % name
    assert set(built_lib.SIGNATURES) | {'mk_last_error'} == declared
    assert built_lib.load().mk_version() >= 100


def test_sass_is_sm100a(built_lib):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_module_api_matches_reference_signatures():
    from modules.generator import MotionTransferGenerator
    from modules.discriminator import Discriminator
    from modules.keypoint_detector import KPDetector
    from modules.dense_motion_module import DenseMotionModule
    from modules.movement_embedding import MovementEmbeddingModule
    want = {
        KPDetector: ['block_expansion', 'num_kp', 'num_channels', 'max_features', 'num_blocks', 'temperature',
                     'kp_variance', 'scale_factor', 'clip_variance'],
        DenseMotionModule: ['block_expansion', 'num_blocks', 'max_features', 'mask_embedding_params', 'num_kp',
                            'num_channels', 'kp_variance', 'use_correction', 'use_mask', 'bg_init',
                            'num_group_blocks', 'scale_factor'],
        MotionTransferGenerator: ['num_channels', 'num_kp', 'kp_variance', 'block_expansion', 'max_features',
                                  'num_blocks', 'num_refinement_blocks', 'dense_motion_params',
                                  'kp_embedding_params', 'interpolation_mode'],
        MovementEmbeddingModule: ['num_kp', 'kp_variance', 'num_channels', 'use_deformed_source_image',
                                  'use_difference', 'use_heatmap', 'add_bg_feature_map', 'heatmap_type', 'norm_const',
                                  'scale_factor'],
        Discriminator: ['num_channels', 'num_kp', 'kp_variance', 'scale_factor', 'block_expansion', 'num_blocks',
                        'max_features', 'kp_embedding_params'],
    }
    for cls, names in want.items():
        assert list(inspect.signature(cls.__init__).parameters)[1:] == names, cls


@pytest.mark.parametrize('name', ['actions', 'bair', 'moving-gif', 'nemo', 'shapes', 'taichi', 'vox', 'vox-full'])
def test_every_config_builds_with_reference_state_dict_keys(name):
    """run.py:50-63 kwargs-splat works for all 8 YAMLs; state_dict keys/shapes equal the oracle's (== reference's)."""
    from modules.generator import MotionTransferGenerator
    from modules.discriminator import Discriminator
    from modules.keypoint_detector import KPDetector
    from oracle import monkey_oracle as mo
    cfg = helpers.load_config(name)
    mp = cfg['model_params']
    assert len(cfg['train_params']['loss_weights']['reconstruction']) == mp['discriminator_params']['num_blocks'] + 1
    with torch.device('meta'):
        prods = (MotionTransferGenerator(**mp['generator_params'], **mp['common_params']),
                 Discriminator(**mp['discriminator_params'], **mp['common_params']),
                 KPDetector(**mp['kp_detector_params'], **mp['common_params']))
        oracles = mo.build_from_config(cfg)
    for p, o in zip(prods, oracles):
        a, b = p.state_dict(), o.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape, k


def test_product_does_not_import_oracle():
    for base in ('modules', 'sync_batchnorm', 'monkey-net_b200'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith('.py'):
                    src = open(os.path.join(dirpath, f)).read()
                    assert 'import oracle' not in src and 'from oracle' not in src, os.path.join(dirpath, f)
