"""Arithmetic of the reference-precision tensor-core mode (csrc/conv_halo.cu), emulated on the CPU: the TF32 main term
plus ONE BF16 MMA for both cross terms must stay at the level of fp32 accumulation itself, far from 1xTF32."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import precision_emul as pe  # noqa: E402


def test_rna_tf32_keeps_ten_mantissa_bits():
    v = torch.tensor([1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, -1.0 - 2.0 ** -11, 3.14159274])
    h = pe.rna_tf32(v)
    assert torch.all((h.view(torch.int32) & 0x1fff) == 0)
    assert float(h[0]) == 1.0 + 2.0 ** -10 and float(h[2]) == -1.0 - 2.0 ** -10    # ties away from zero
    assert float((v - h).abs().max()) <= 2.0 ** -11 * 4


def test_cross_term_scheme_is_fp32_level():
    torch.manual_seed(1)
    for K in (256, 9 * 48):
        A, B = torch.randn(2048, K), torch.randn(K, 48) * 0.05
        e = pe.errors(A, B)
        assert e['tf32'][1] > 1e-4                                  # what the split has to remove
        assert e['tf32+bf16cross'][1] < 1.5e-6                      # 2^-20-level operand error
        assert e['tf32+bf16cross'][1] < 4 * e['fp32 matmul'][1]     # same order as fp32 accumulation
        assert e['tf32+bf16cross'][0] < 5e-6                        # max error, relative to the output's max


def test_cross_term_scheme_on_activation_like_data():
    torch.manual_seed(2)
    A, B = torch.rand(2048, 9 * 64), torch.rand(9 * 64, 32) * 0.05     # post-ReLU-like: aligned sums
    e = pe.errors(A, B)
    assert e['tf32+bf16cross'][1] < e['fp32 matmul'][1]
    assert e['tf32+bf16cross'][0] < 1e-6
