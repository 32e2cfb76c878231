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


# The product default is the tensor-core (TF32) convolution; the tight fp32-tolerance parity suites pin the exact
# FFMA convolution, tests/test_gpu_3_tc.py covers the tensor-core mode at the north-star tolerances.
_FP32_MODULES = ('test_gpu_1_ops', 'test_gpu_2_modules', 'test_gpu_4_graph')


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
