import pytest
import torch

from helpers import make_manywell, make_gmm, make_aldp, rel_err, assert_parity
import fab_torch_b200 as fb
from oracle.targets import OracleDiagGaussian

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [2, 32, 128])
def test_manywell(dim):
    to, tp = make_manywell(dim)
    torch.manual_seed(100 + dim)          # (was unseeded: the outcome depended on the test order)
    x = torch.randn(257, dim) * 1.7
    x64 = x.double().requires_grad_(True)
    ref = to.log_prob(x64)
    gref = torch.autograd.grad(ref.sum(), x64)[0]
    xg = x.cuda().requires_grad_(True)
    lp = tp.log_prob(xg)
    g = torch.autograd.grad(lp.sum(), xg)[0]
    x32 = x.clone().requires_grad_(True)
    ref32 = to.log_prob(x32)
    g32 = torch.autograd.grad(ref32.sum(), x32)[0]
    assert_parity(lp, ref, ref32, "log_p")
    assert_parity(g, gref, g32, "grad_log_p")
    assert abs(float(fb.ManyWellEnergy(32).log_Z) - 164.69567532) < 1e-6


@pytest.mark.parametrize("dim,n_mixes,loc", [(2, 40, 40.0), (2, 4, 8.0), (3, 33, 5.0)])
def test_gmm(dim, n_mixes, loc):
    to64, _, tp = make_gmm(dim, n_mixes, loc, 1.0)
    x = (torch.rand(300, dim) - 0.5) * 2.2 * loc
    x64 = x.double().requires_grad_(True)
    ref = to64.log_prob(x64)
    gref = torch.autograd.grad(ref.sum(), x64)[0]
    xg = x.cuda().requires_grad_(True)
    lp = tp.log_prob(xg)
    g = torch.autograd.grad(lp.sum(), xg)[0]
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(lp).cpu(), fin)
    assert rel_err(lp.cpu()[fin], ref[fin]) < 1e-5
    assert rel_err(g.cpu()[fin], gref[fin]) < 1e-4


def test_gmm_mask_and_gaussian_target():
    _, _, tp = make_gmm(2, 4, 8.0)
    far = torch.full((3, 2), 1.0e3).cuda()
    assert torch.isinf(tp.log_prob(far)).all()          # gmm.py:63-65
    loc = torch.tensor([-0.5, -0.5, -0.5])
    tgt = fb.DiagGaussianTarget(loc, 1.0)
    ref = OracleDiagGaussian(loc.double(), 1.0)
    x = torch.randn(100, 3)
    assert rel_err(tgt.log_prob(x.cuda()), ref.log_prob(x.double())) < 1e-5


@pytest.mark.parametrize("dim", [60, 6])
def test_aldp_surrogate(dim):
    """BASELINE config 5 target: value and closed-form gradient vs the fp64 oracle (autograd); the
    default tables (no `tables=`) must be the same pure function of (dim, seed)."""
    to64, to, tp = make_aldp(dim)
    assert torch.equal(fb.AldpSurrogateEnergy(dim, 0).p0.cpu(), tp.p0.cpu())
    x = torch.randn(300, dim) * 1.5
    x64 = x.double().requires_grad_(True)
    ref = to64.log_prob(x64)
    gref = torch.autograd.grad(ref.sum(), x64)[0]
    x32 = x.clone().requires_grad_(True)
    ref32 = to.log_prob(x32)
    g32 = torch.autograd.grad(ref32.sum(), x32)[0]
    xg = x.cuda().requires_grad_(True)
    lp = tp.log_prob(xg)
    g = torch.autograd.grad(lp.sum(), xg)[0]
    assert_parity(lp, ref, ref32, "log_p")
    assert_parity(g, gref, g32, "grad_log_p", floor=2e-5)
