import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


# Module-, step- and graph-level suites run in the product's DEFAULT arithmetic ('auto': reference-precision tf32x3 tensor-core convs
# while autograd records and in the keypoint detector, 1xTF32 for no_grad inference) - the mode bench.py times.
# Only the op-level suite pins 'fp32': its convolution cases are the tests of the exact FFMA kernels themselves.
_FP32_MODULES = ('test_gpu_1_ops',)


@pytest.fixture(autouse=True)
def _pin_conv_mode(request):
    name = request.module.__name__.rsplit('.', 1)[-1]
    if name not in _FP32_MODULES:
        yield
        return
    from monkey_net_b200 import ops
    prev = ops.CONV_MODE
    ops.set_conv_mode('fp32')
    try:
        yield
    finally:
        ops.set_conv_mode(prev)
