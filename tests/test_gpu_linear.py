"""Dense layer (com_dense) through iisan_linear_forward / _backward: exact mode (fp32 FMA) and fast mode (TF32 mma.sync tiles)
against torch.nn.functional.linear on the same inputs (ragged row counts included)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("bf16", 2e-3)])
@pytest.mark.parametrize("rows,n,k", [(5632, 64, 192), (100, 64, 192), (33, 16, 64), (1, 64, 192)])
def test_linear_matches_torch(mode, tol, rows, n, k):
    from iisan_b200.ops import LinearFn
    from iisan_b200.precision import compute_mode, set_compute_mode
    g = torch.Generator(device="cuda").manual_seed(rows + n)
    x = torch.randn(rows, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.1
    b = torch.randn(n, device="cuda", generator=g)
    gy = torch.randn(rows, n, device="cuda", generator=g)
    xs, ws, bs = (t.clone().requires_grad_(True) for t in (x, w, b))
    set_compute_mode(mode)
    try:
        y = LinearFn.apply(xs, ws, bs, compute_mode())
        y.backward(gy)
    finally:
        set_compute_mode(None)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.linear(xr, wr, br)
    ref.backward(gy)
    rel = lambda a, c: float((a - c).norm() / (c.norm() + 1e-12))
    assert rel(y, ref) <= tol
    assert rel(xs.grad, xr.grad) <= tol
    assert rel(ws.grad, wr.grad) <= tol
    assert rel(bs.grad, br.grad) <= 1e-5
