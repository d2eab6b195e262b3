"""CPU emulation of the operand format the tensor-core GEMMs use for fp32 parity ("bf16x3"; csrc/gemm_tc.cu,
caphn_split_bf16): x = hi + lo with hi = rn_bf16(x), lo = rn_bf16(x - hi); every k-slice contributes hi*hi + hi*lo + lo*hi
(the lo*lo term is dropped) into an fp32 accumulator.  No GPU: this pins the error budget of the FORMAT -- what the scheme
can reach with an ideal fp32 accumulator -- against the 1e-4 logits tolerance of BASELINE.json; the GPU tests measure the
kernel itself (TMEM accumulation adds its own truncation, tests/test_gpu_gemm_tc.py)."""
import pytest
import torch


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.float(), lo.float()


def bf16x3(a, b):
    ah, al = split(a)
    bh, bl = split(b)
    return ah @ bh.T + ah @ bl.T + al @ bh.T


@pytest.mark.parametrize("M,N,K", [(64, 96, 150), (64, 96, 200), (32, 48, 2048), (16, 24, 10240)])
def test_split_format_reaches_fp32_class_accuracy(M, N, K):
    g = torch.Generator().manual_seed(K)
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    ref = a.double() @ b.double().T
    scale = ref.abs().max().item()
    err3 = (bf16x3(a, b).double() - ref).abs().max().item() / scale
    err1 = ((a.to(torch.bfloat16).float() @ b.to(torch.bfloat16).float().T).double() - ref).abs().max().item() / scale
    errf = ((a @ b.T).double() - ref).abs().max().item() / scale
    # hi + lo carries 16 mantissa bits per operand and the lo*lo term is dropped: the format's own floor is ~5e-6 of
    # max|C| whatever K is (the GPU kernel measures 5e-6 at K = 200: it sits ON this floor) -- an order of magnitude above
    # plain fp32, more than an order inside the 1e-4 budget; a single bf16 product per k-slice (2-3e-3) is far outside it
    assert err3 < 1e-5, err3
    assert err3 < 40 * max(errf, 1e-7)
    assert err1 > 1e-4 or K < 200, (err1, "plain bf16 operands would already meet the budget: the split would be pointless")
    assert err1 > 50 * err3


def test_split_is_exact_to_sixteen_bits_and_padding_is_neutral():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1000, generator=g) * torch.logspace(-6, 6, 1000)
    hi, lo = split(x)
    rel = ((hi + lo) - x).abs() / x.abs()
    assert rel.max().item() < 2.0 ** -16            # two bf16 mantissas (8 + 8 bits, hidden bits included) cover 16 bits
    # zero padding of K to a multiple of 64 (Kp) contributes nothing: hi(0) = lo(0) = 0
    z = torch.zeros(7)
    assert split(z)[0].abs().sum() == 0 and split(z)[1].abs().sum() == 0
