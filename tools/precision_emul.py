"""CPU emulation of the reference-precision operand split of the tensor-core kernels (csrc/conv_halo.cu):

    a * b  ~=  tf32(a) * tf32(b)                                   one kind::tf32 MMA, K = 8
             + bf16(a - tf32(a)) * bf16(b) + bf16(a) * bf16(b - tf32(b))    one kind::f16 BF16 MMA, K = 16

against 1xTF32, 3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) and torch's fp32 matmul, all measured against float64.
The products are accumulated in float64 here, i.e. the numbers isolate the OPERAND rounding of each scheme.

    python tools/precision_emul.py
"""
import torch


def rna_tf32(v):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 mantissa bits kept"""
    i = v.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def bf16(v):
    return v.bfloat16().float()


def schemes(A, B):
    """A: (M, K), B: (K, N) fp32 -> dict name -> float64 result"""
    d = lambda t: t.double()
    Ah, Bh = rna_tf32(A), rna_tf32(B)
    one = d(Ah) @ d(Bh)
    x3 = one + d(rna_tf32(A - Ah)) @ d(Bh) + d(Ah) @ d(rna_tf32(B - Bh))
    hyb = one + d(bf16(A - Ah)) @ d(bf16(B)) + d(bf16(A)) @ d(bf16(B - Bh))
    return {'tf32': one, '3xtf32': x3, 'tf32+bf16cross': hyb, 'fp32 matmul': d(A @ B), 'exact': d(A) @ d(B)}


def errors(A, B):
    r = schemes(A, B)
    ref = r.pop('exact')
    rms = ref.pow(2).mean().sqrt()
    return {k: (float((v - ref).abs().max() / ref.abs().max()), float((v - ref).pow(2).mean().sqrt() / rms))
            for k, v in r.items()}


def main():
    torch.manual_seed(0)
    for K in (16 * 16, 9 * 48, 9 * 512):
        for name, A, B in (('randn', torch.randn(4096, K), torch.randn(K, 64) * 0.05),
                           ('positive', torch.rand(4096, K), torch.rand(K, 64) * 0.05)):
            for k, (mx, rms) in errors(A, B).items():
                print('%-8s K=%-5d %-16s max/|max| %.2e   rms/rms %.2e' % (name, K, k, mx, rms))


if __name__ == '__main__':
    main()
