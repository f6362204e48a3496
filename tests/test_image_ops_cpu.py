"""CPU: the bicubic-antialias tap tables (forward and adjoint) reproduce aten's upsample_bicubic2d_aa exactly."""
import pytest
import torch
import torch.nn.functional as F

from comat_b200 import image_ops as IO


@pytest.mark.parametrize("i,o", [(510, 384), (254, 384), (190, 384), (64, 24), (33, 33)])
def test_dense_operator_matches_aten(i, o):
    torch.manual_seed(i)
    x = torch.rand(2, 3, i, i, dtype=torch.float64)
    M = IO.dense_operator(i, o).double()
    got = M @ x @ M.t()
    ref = F.interpolate(x, size=(o, o), mode="bicubic", antialias=True, align_corners=False)
    assert (got - ref).abs().max() < 2e-6
    # adjoint tables == transpose of the forward operator
    st, ct, w = IO.aa_bicubic_taps_transposed(i, o)
    Mt = torch.zeros(i, o, dtype=torch.float64)
    for j in range(i):
        Mt[j, int(st[j]):int(st[j]) + int(ct[j])] = w[j, :int(ct[j])].double()
    assert (Mt - M.t()).abs().max() < 1e-7
