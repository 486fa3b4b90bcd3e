"""the reference's own training / inference loops run on the drop-in modules.

INTEGRATION.md section 2 claims that `train()` / `infer()` of the reference (src/run.py:96-319, :15-50) run unchanged once
its imports point at tinynerf_b200.  /root/reference is not on the GPU box, so the loop bodies are restated here with the
reference's own idiom, statement for statement where it touches the path: a torch DataLoader of dict batches, the dynamic
batch accumulator with its IN-PLACE `info[:, 0] += current_size` and `torch.cat` (src/run.py:215-244), stock
`torch.optim.Adam` + `MultiStepLR` + `GradScaler(2**10)` that is never unscaled (:186-201,258-261), `occupancy_grid.update`
with a lambda over the two modules (:248-249), `loss.detach().cpu().item()` and `occupancy_grid.occupancy()` (:263-264),
`state_dict()` (:308), and the chunked render with `.cpu()` per chunk (:34-44).  None of the trainer conveniences of
tinynerf_b200.run are used.  The first iteration's loss is also held against the reference restatement on the same batch."""
import math

import pytest
import torch
from torch.utils.data import DataLoader, Dataset

from oracle import ref_port as rp
from tinynerf_b200 import synthetic
from tinynerf_b200.core import (ContractionAABB, ContractionMip360, NerfRenderer, OccupancyGrid, RayMarcherAABB,
                                RayMarcherUnbounded, RayProvider)
from tinynerf_b200.models import (CobafaFeatureField, KPlanesFeatureField, VanillaColorDecoder, VanillaFeatureMLP,
                                  VanillaOpacityDecoder)

pytestmark = pytest.mark.gpu


class Rays(Dataset):   # RaysDataset.__getitem__ (src/data.py:102-120)
    def __init__(self, n, seed):
        self.rays_o, self.rays_d = synthetic.blender_rays(n, seed=seed)
        self.rgbs = (0.5 + 0.5 * torch.nn.functional.normalize(self.rays_d, dim=-1)).clamp(0, 1)

    def __len__(self):
        return self.rays_o.size(0)

    def __getitem__(self, idx):
        return {"rays_o": self.rays_o[idx], "rays_d": self.rays_d[idx], "rgbs": self.rgbs[idx]}


@pytest.mark.parametrize("method,scene_type", [("kplanes", "aabb"), ("cobafa", "aabb"), ("kplanes", "unbounded"), ("vanilla", "aabb")])
def test_reference_train_and_infer_loops_run_on_the_dropin_modules(method, scene_type):
    device = torch.device("cuda")
    torch.manual_seed(3)
    batch_size, n_samples = 256, 64
    train_rays = Rays(1 << 13, seed=5)
    train_loader = DataLoader(train_rays, batch_size=batch_size, shuffle=True, num_workers=0, pin_memory=True)
    # ---- model construction, src/run.py:104-184 ----
    bs_ratio = 4096 / batch_size
    steps = int(2048 * bs_ratio)
    occupancy_grid_updates = 2   # int(16 * bs_ratio) in the reference; small here so the update runs inside the test
    occupancy_grid_threshold, occupancy_grid_res = 0.01, 128
    occupancy_grid_decay = occupancy_grid_threshold ** (1 / 16)
    tv_reg_alpha, l1_reg_alpha = 0.0001, 0.0
    if method == "vanilla":
        feature_module = VanillaFeatureMLP(10, 256, 8)
    elif method == "kplanes":
        feature_module = KPlanesFeatureField(32)
    else:
        feature_module = CobafaFeatureField(basis_res=torch.linspace(32.0, 128, 6).int().tolist(), coef_res=64,
                                            freqs=torch.linspace(2.0, 8.0, 6).tolist(), channels=[8, 8, 8, 4, 4, 4], mlp_hidden_dim=128)
    dim = feature_module.feature_dim
    sigma_decoder = VanillaOpacityDecoder(dim)
    rgb_decoder = VanillaColorDecoder(8, dim, 64, 3)
    if scene_type == "unbounded":
        ray_marcher = RayMarcherUnbounded(n_samples, 0.1, 1e5, uniform_range=1.3)
        contraction = ContractionMip360(order=float("inf"))
    else:
        aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]]).to(device)
        ray_marcher = RayMarcherAABB(aabb, n_samples, 0.1)
        contraction = ContractionAABB(aabb)
    occupancy_grid = OccupancyGrid(size=occupancy_grid_res, step_size=ray_marcher.step_size, threshold=occupancy_grid_threshold,
                                   decay=occupancy_grid_decay).to(device)
    ray_provider = RayProvider(occupancy_grid, contraction, ray_marcher)
    renderer = NerfRenderer(feature_module, sigma_decoder, rgb_decoder, bg_color=torch.ones(3)).to(device)
    optimizer = torch.optim.Adam(renderer.parameters(), lr=1e-2, eps=1e-15, weight_decay=1e-5)
    scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[steps // 2, steps * 3 // 4, steps * 5 // 6, steps * 9 // 10], gamma=0.33)
    scaler = torch.amp.GradScaler("cuda", init_scale=2 ** 10)
    loss_fn = torch.nn.MSELoss()

    # ---- the loop, src/run.py:213-264 ----
    train_iter = iter(train_loader)
    target_sample_size = batch_size * n_samples
    losses = []
    for train_step in range(5):
        with torch.no_grad():
            current_size, projected_size, tmp_count = 0, 0, 0
            acc_info, acc_samples, acc_rgbs = [], [], []
            while projected_size < target_sample_size:
                try:
                    data = next(train_iter)
                except StopIteration:
                    train_iter = iter(train_loader)
                    data = next(train_iter)
                rays_o, rays_d, rgbs = data["rays_o"].to(device), data["rays_d"].to(device), data["rgbs"].to(device)
                samples, info = ray_provider(rays_o, rays_d, training=True)
                info[:, 0] += current_size          # in place, as the reference does (voids the trusted-partition tag)
                acc_info.append(info); acc_samples.append(samples); acc_rgbs.append(rgbs)
                current_size += samples.size(0)
                tmp_count += 1
                projected_size = int(current_size * (1 + 1 / tmp_count))
            packed_samples, packed_rgbs, packing_info = torch.cat(acc_samples, 0), torch.cat(acc_rgbs, 0), torch.cat(acc_info, 0)
        renderer.train()
        if train_step % occupancy_grid_updates == 0:
            occupancy_grid.update(lambda t: renderer.sigma_decoder(renderer.feature_module(t)))
        rendered_rgbs = renderer(packed_samples, packing_info)
        loss = loss_fn(rendered_rgbs, packed_rgbs)
        if train_step == 0 and method == "kplanes":   # the same batch through the reference restatement
            planes = [[p.plane for p in s] for s in feature_module.planes]
            s_l = [(l.weight, l.bias) for l in sigma_decoder.net.linears()]
            c_l = [(l.weight, l.bias) for l in rgb_decoder.net.linears()]
            with torch.no_grad():
                want = rp.render(lambda x: rp.kplanes_features(planes, x), lambda f: rp.sigma_head(s_l, f),
                                 lambda f, dd: rp.rgb_head(c_l, 8, f, dd), packed_samples, packing_info, torch.ones(3))
            assert float(loss) == pytest.approx(float(loss_fn(want, packed_rgbs)), rel=1e-5)
        if method == "kplanes":
            loss += renderer.feature_module.loss_tv() * tv_reg_alpha
            loss += renderer.feature_module.loss_l1() * l1_reg_alpha
        optimizer.zero_grad()
        scaler.scale(loss).backward()
        optimizer.step()
        scheduler.step()
        losses.append(loss.detach().cpu().item())
        occ = occupancy_grid.occupancy()
        assert 0.0 < occ <= 1.0 and packed_samples.size(0) > 0
    assert all(math.isfinite(l) for l in losses)
    assert all(torch.isfinite(p).all() for p in renderer.parameters())
    sd = renderer.state_dict()      # src/run.py:308
    assert any(k.startswith("feature_module.") for k in sd) and any(k.startswith("rgb_decoder.net.net") for k in sd)

    # ---- infer, src/run.py:25-45 ----
    renderer.eval()
    o, d = synthetic.camera_rays(40, 30, 0.5 * 40 / math.tan(0.5 * 0.6911112), [2.6, -1.9, 2.4])
    rendered = []
    with torch.no_grad():
        for k in range(0, len(o), 512):
            samples, info = ray_provider(o[k:k + 512].to(device), d[k:k + 512].to(device), training=False)
            rendered.append(renderer(samples, info).cpu())
    img = torch.cat(rendered, dim=0).view(30, 40, 3)
    assert torch.isfinite(img).all() and float(img.min()) >= 0.0 and float(img.max()) <= 1.0 + 1e-6
    assert (255.0 * img).type(torch.uint8).numpy().shape == (30, 40, 3)
