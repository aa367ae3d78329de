"""The tcgen05/TMEM/TMA GEMM building block (msf_gemm_bf16) against a plain
PyTorch fp32 reference of the same contraction on the same bf16-rounded operands.
Tolerance: fp32 accumulate of exact bf16 products, so only summation order
differs -> max-abs <= 1e-3 * sqrt(K) scale for fp32 output; bf16 output adds one rounding."""
import importlib

import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.gpu


def _ops():
    return importlib.import_module(load_pkg().__name__ + ".ops")


def _rand(shape, seed, ld=None):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(*shape, generator=g)
    if ld is not None:  # padded row pitch
        buf = torch.zeros(shape[0], ld)
        buf[:, :shape[1]] = t
        return buf.cuda().to(torch.bfloat16)[:, :shape[1]]
    return t.cuda().to(torch.bfloat16)


K_MAJOR = [  # m, n, k
    (128, 256, 64), (4096, 256, 128), (4096, 256, 256), (300, 200, 72), (1000, 25, 256), (77, 128, 256),
    (4096, 128, 256), (129, 64, 8), (4096, 512, 256),
]


@pytest.mark.parametrize("m,n,k", K_MAJOR)
@pytest.mark.parametrize("out", [torch.float32, torch.bfloat16])
def test_k_major_matches_fp32_reference(m, n, k, out):
    ops = _ops()
    a, b = _rand((m, k), 1), _rand((n, k), 2)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(3)).cuda()
    ref = a.float() @ b.float().T + bias
    got = ops.gemm_bf16(a, b, bias=bias, out_dtype=out)
    tol = 2e-4 * k ** 0.5 if out == torch.float32 else 2e-2 * max(1.0, float(ref.abs().max()) / 4)
    assert float((got.float() - ref).abs().max()) <= tol
    got_relu = ops.gemm_bf16(a, b, bias=bias, relu=True, out_dtype=out)
    assert float((got_relu.float() - ref.clamp_min(0)).abs().max()) <= tol


MN_MAJOR = [  # k (contraction = windows), m, n
    (64, 64, 64), (4096, 256, 256), (4096, 256, 128), (1000, 256, 256), (4096, 32, 256), (333, 128, 64),
    (8192, 256, 128),
    # CTA-pair kernel (wg2_gemm.cu, n % 128 == 0): several 256 x 256 tiles, ragged m, a 128-wide last column tile,
    # contraction of 3 / 1 k-blocks (uneven halves / no split)
    (4096, 512, 512), (4096, 704, 384), (130, 256, 128), (64, 256, 256), (2048, 25, 256),
    (512, 2560, 2560),   # 100 tiles on 74 CTA pairs: no split, clusters loop over several tiles
]


@pytest.mark.parametrize("k,m,n", MN_MAJOR)
def test_mn_major_matches_fp32_reference(k, m, n):
    ops = _ops()
    lda = 32 if m < 64 else None
    a, b = _rand((k, m), 4, ld=lda), _rand((k, n), 5)
    ref = a.float().T @ b.float()
    got = ops.gemm_bf16(a, b, mn_major=True)
    assert got.shape == (m, n)
    assert float((got - ref).abs().max()) <= 2e-4 * k ** 0.5 * max(1.0, float(ref.abs().max()) / 64)


def test_split_contraction_is_bit_reproducible():
    """The two halves of a split contraction meet in the destination in either order (wg2_gemm.cu): same bits every run."""
    ops = _ops()
    a, b = _rand((4096, 256), 6), _rand((4096, 256), 7)
    first = ops.gemm_bf16(a, b, mn_major=True).clone()
    for _ in range(20):
        assert torch.equal(ops.gemm_bf16(a, b, mn_major=True), first)


def test_rejects_misaligned_operands():
    ops, pkg = _ops(), load_pkg()
    a, b = _rand((128, 20), 1), _rand((64, 20), 2)  # ld = 20 elements: rows not 16-byte aligned
    with pytest.raises(pkg.MsfError):
        ops.gemm_bf16(a, b)
