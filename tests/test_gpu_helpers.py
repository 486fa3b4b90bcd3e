"""The stand-alone helper callables of the mirror (contractions, marchers, single-plane / single-grid lookups, L1 / TV
regularisers, positional encoding) run their own kernels (csrc/helpers.cu) -- checked here against the reference's
golden vectors and the PyTorch restatement on the same GPU.  Bit-exact where the reference's arithmetic is pinned
(march / contraction / PE), 1e-5 relative for interpolation and reductions."""
import pytest
import torch

from oracle import ref_port as rp
from tinynerf_b200 import core, models

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_marchers_and_contractions_match_reference_golden(golden):
    g = golden("provider_aabb")
    m = core.RayMarcherAABB(g["aabb"].to(DEV), 64, 0.1)
    t, s = m(g["rays_o"].to(DEV), g["rays_d"].to(DEV))
    assert torch.equal(t.cpu(), g["t_values"]) and torch.equal(s.cpu(), g["step_sizes"])
    g = golden("provider_unbounded")
    m = core.RayMarcherUnbounded(64, 0.1, 1e5, uniform_range=1.7)
    t, s = m(g["rays_o"].to(DEV), g["rays_d"].to(DEV))
    # the tables are built with the reference's op sequence on the GPU: torch's CUDA linspace / reciprocal may differ from
    # the CPU golden in the last bit
    assert torch.allclose(t.contiguous().cpu(), g["t_values"], rtol=2e-7, atol=0)
    assert torch.allclose(s.contiguous().cpu(), g["step_sizes"], rtol=1e-5, atol=1e-9)
    pts = (torch.randn(4096, 3, generator=torch.Generator().manual_seed(0)) * 3).to(DEV)
    c, mask = core.ContractionMip360()(pts)
    assert mask is None and torch.equal(c, rp.contract_mip360(pts))
    aabb = torch.tensor([[0.0, 0, 0], [2.0, 2.5, 3]], device=DEV)
    c, mask = core.ContractionAABB(aabb)(pts.view(64, 64, 3))
    c_ref, m_ref = rp.contract_aabb(pts.view(64, 64, 3), aabb)
    assert c.shape == c_ref.shape and torch.equal(c, c_ref) and torch.equal(mask, m_ref)


def test_single_plane_and_grid_lookups_vs_grid_sample():
    torch.manual_seed(3)
    plane = models.KPlanesFeaturePlane(8, (48, 80)).to(DEV)
    x = (torch.rand(5000, 2, device=DEV) * 2.2 - 1.1)
    out = plane(x)
    ref = rp.plane_lookup(plane.plane, x)
    assert out.shape == ref.shape and torch.allclose(out, ref, rtol=1e-5, atol=1e-7)
    g = torch.randn_like(out)
    (out * g).sum().backward()
    mine = plane.plane.grad.clone()
    plane.plane.grad = None
    (rp.plane_lookup(plane.plane, x) * g).sum().backward()
    assert torch.allclose(mine, plane.plane.grad, rtol=1e-5, atol=1e-6 * float(plane.plane.grad.abs().max()))
    grid = models.CobafaGrid((12, 20, 28), 4).to(DEV)
    x3 = (torch.rand(5000, 3, device=DEV) * 2.2 - 1.1)
    out = grid(x3)
    ref = rp.grid3_lookup(grid.grid, x3)
    assert out.shape == ref.shape and torch.allclose(out, ref, rtol=1e-5, atol=1e-7)
    g = torch.randn_like(out)
    (out * g).sum().backward()
    mine = grid.grid.grad.clone()
    grid.grid.grad = None
    (rp.grid3_lookup(grid.grid, x3) * g).sum().backward()
    assert torch.allclose(mine, grid.grid.grad, rtol=1e-5, atol=1e-6 * float(grid.grid.grad.abs().max()))
    # leading batch dimensions survive
    assert plane(x.view(50, 100, 2)).shape == (50, 100, 8) and grid(x3.view(10, 500, 3)).shape == (10, 500, 4)


def test_regularisers_and_encoding_match_reference_golden(golden):
    g = golden("kplanes")
    torch.manual_seed(21)
    f = models.KPlanesFeatureField(32).to(DEV)
    assert float(f.loss_tv()) == pytest.approx(g["tv"], rel=1e-5)
    l1 = f.loss_l1()
    assert float(l1) == pytest.approx(g["l1"], rel=1e-6)
    (l1 * 2.0).backward()
    p = f.planes[2][1].plane
    assert torch.equal(p.grad, torch.sign(p.detach()) * (2.0 / 9 / p.numel()))
    single = f.planes[0][0]
    assert float(single.loss_l1()) == pytest.approx(float(single.plane.detach().abs().mean()), rel=1e-6)
    g = golden("heads")
    pe = models.PositionalEncoding(8).to(DEV)
    assert torch.allclose(pe(g["dirs"].to(DEV)).cpu(), g["pe"], rtol=0, atol=2e-6)   # CPU sinf/cosf of the golden run vs GPU
    assert torch.equal(pe(g["dirs"].to(DEV)), rp.positional_encoding(g["dirs"].to(DEV), 8))   # torch's CUDA sin/cos: bit-exact
    x = torch.randn(7, 5, 3, device=DEV)
    assert torch.equal(models.PositionalEncoding(10).to(DEV)(x), rp.positional_encoding(x, 10))


def test_every_reference_mlp_shape_runs_on_the_tensor_core_kernels():
    """a17: no MLP of the reference's three methods is sent to cuBLAS; what the kernels do not cover raises."""
    fm = models.VanillaFeatureMLP(10, 256, 8).to(DEV)
    assert fm(torch.rand(300, 3, device=DEV)).shape == (300, 256)
    odd = models.MLP(16, 256, 1, 3).to(DEV)   # a 3-wide head behind a 256-wide hidden layer: not covered -> loud
    with pytest.raises(NotImplementedError, match="no cuBLAS fallback"):
        odd(torch.rand(8, 16, device=DEV))
