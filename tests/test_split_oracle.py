"""The split-precision arithmetic of the tcgen05 Linear kernels, pinned on the CPU (oracle/allset_oracle.py::linear_split):
three bf16 terms per operand and the six products down to 2^-18 reproduce an fp32 Linear (reference src/layers.py:575) to
fp32 accuracy; two terms / three products -- the usual "3x" split -- are an order of magnitude coarser, which is what the
fp64-arbiter diagnostic on the GPU (scripts/diag_split.py, profiles/r02_split_diag_*) found to flip ReLUs."""
import torch

import allset_oracle as O


def _case(rows=4096, d=128, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, d, generator=g) * 1.5 + 0.25
    w = (torch.rand(d, d, generator=g) * 2 - 1) / d ** 0.5
    return x, w


def test_bf16_terms_reconstruct_the_operand():
    x, _ = _case()
    t = O.bf16_terms(x, 3)
    assert all(torch.equal(ti, ti.to(torch.bfloat16).float()) for ti in t)          # every term is a bf16 number
    assert (x - (t[0] + t[1] + t[2])).abs().max().item() <= 2.0 ** -23 * x.abs().max().item()
    assert (x - (t[0] + t[1])).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()


def test_three_term_split_matches_an_fp32_linear_and_two_terms_do_not():
    x, w = _case()
    ref = x.double() @ w.double().t()
    scale = ref.abs().max().item()
    e_fp32 = (x @ w.t() - ref).abs().max().item() / scale
    e3 = (O.linear_split(x, w, 3).double() - ref).abs().max().item() / scale
    e2 = (O.linear_split(x, w, 2).double() - ref).abs().max().item() / scale
    assert e3 <= 2e-6 and e3 <= 4 * max(e_fp32, 1e-7)        # the fp32 class: what tests/test_linear_tc.py holds the kernel to
    assert e2 >= 3 * e3                                      # the "3x" split is visibly coarser ...
    assert e2 <= 1e-4                                        # ... though still inside the 1e-4 bar of the forward logits


def test_split_linear_is_linear_and_handles_ragged_shapes():
    x, w = _case(rows=77, d=64, seed=3)
    a = O.linear_split(x, w, 3)
    assert a.shape == (77, 64)
    b = O.linear_split(2 * x, w, 3)                          # scaling by a power of two commutes with the bf16 rounding
    assert torch.equal(b, 2 * a)
