#!/usr/bin/env python
"""bench.py -- ray samples/sec per train step of the packed-ray hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--no-extras]

Workloads (config.workload; BASELINE.json `configs` 1-5 + the render half):
  kplanes_aabb_2e18        config 2 (DEFAULT, the configuration the metric is quoted on): K-Planes + vanilla heads on a
                           synthetic 800x800 Blender-shaped scene, AABB +-1.5, 128^3 occupancy grid held at a seeded
                           analytic state (ball + torus + floaters; the cadence update IS executed, then the state is
                           restored), dynamic batches of 1024-ray chunks x 256 samples -> ~2^18 packed samples per step per GPU.
  cobafa_aabb_dyn          config 3: Cobafa field (Dropout 0.01 active) on the same scene, dynamic batches, rays sharded.
  kplanes_unbounded_decay  config 4: K-Planes on a COLMAP-shaped unbounded scene (Mip-360 contraction), the grid starts
                           all-ones and is NOT pinned: >= 3 occupancy update/decay cycles fall inside the timed window.
  vanilla_dummy            config 1's model (VanillaFeatureMLP(10,256,8), 64 samples/ray) on a tests/dummy-shaped scene
                           (2 cameras x 200x200 = 80,000 rays); its CPU leg is `--impl reference --workload vanilla_dummy`.
  weights_micro            config 5: packed weights fwd+bwd, 2^18..2^26 samples, every rank on its own inputs, aggregate GB/s,
                           beside the UNMODIFIED reference kernel (oracle/_ref/_cuda.so) on the same inputs (R <= 2^20 rays).
  render_800               the render half (src/run.py:15-50): whole 800x800 poses, rays/s.
One "step" = march+pack -> feature gather -> heads -> weights -> composite -> MSE(+TV) -> backward -> Adam (+ the occupancy
update when the reference's cadence says so).  Rays shard across ranks (weak scaling).

value       : packed samples processed by all ranks / max-over-ranks device time, ray store in HBM
e2e         : same through the public API with the ray store in pinned HOST memory (per-batch host gather + H2D inside
              the timed region) and the loss read back (D2H) every step
roofline    : the dominant kernel of ours in the step, CUDA-event timed per launch in the timed region
cpu_baseline: the oracle's PyTorch-CPU restatement of the same step on this box's host cores (rank 0, N=1)
gpu_reference: informational -- the same restatement on the GPU (stock grid_sample / Linear / index_add_ + the reference's
              own weights kernel): what the reference's code does on this B200
parity      : achieved worst-case errors of the CUDA path vs that restatement (oracle/parity.py, the checker)
workloads   : short runs of the other configurations (skipped with --no-extras)
--impl reference: the CPU path alone, as the reference arm, at the SAME per-step sample target.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
import traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ray samples/sec per train step"
UNIT = "samples/s"
N_STORE = 1 << 21          # rays in the synthetic scene store (x36 B = 75 MB; > one epoch of the run)
BATCH, N_SAMPLES = 1024, 256
SEED = 1234

DP_MODE_USED = [os.environ.get("TNF_DP_MODE", "peer")]

TRAIN_SPECS = {
    "kplanes_aabb_2e18": dict(method="kplanes", scene="aabb", n_samples=256, pin_grid=True, rays="blender", min_steps=0),
    "cobafa_aabb_dyn": dict(method="cobafa", scene="aabb", n_samples=256, pin_grid=True, rays="blender", min_steps=0),
    "kplanes_unbounded_decay": dict(method="kplanes", scene="unbounded", n_samples=256, pin_grid=False, rays="colmap", min_steps=200),
    "vanilla_dummy": dict(method="vanilla", scene="aabb", n_samples=64, pin_grid=False, rays="dummy", min_steps=0),
}
WORKLOADS = list(TRAIN_SPECS) + ["weights_micro", "render_800"]


def measured_peaks():
    """(HBM GB/s, dense TF32 TFLOP/s = half the measured bf16 burst figure, source)"""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]) / 2, "measured"
    return 6650.0, 1590.0 / 2, "fallback"


def set_blocking_sync(device_index: int) -> bool:
    """cudaDeviceScheduleBlockingSync for this process's device, before its context exists: a host thread waiting for the
    GPU sleeps instead of spinning.  With one process per GPU plus NCCL's proxy threads, eight spinning main threads on a
    16-core host preempt each other for whole timeslices (we measured 30-60 ms stalls in the per-step-synchronised arm)."""
    try:
        import ctypes
        import glob
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        rt = ctypes.CDLL(cands[0] if cands else "libcudart.so")
        return rt.cudaSetDevice(device_index) == 0 and rt.cudaSetDeviceFlags(4) == 0  # cudaDeviceScheduleBlockingSync
    except Exception:
        return False


# ---- synthetic scenes ---------------------------------------------------------------
def make_scene(kind: str, n_rays: int, seed: int):
    """-> (rays_o, rays_d, rgbs, scene_scale).  Colours are uniform random: throughput depends only on the occupancy."""
    from tinynerf_b200 import synthetic
    scale = 1.0
    if kind == "blender":
        o, d = synthetic.blender_rays(n_rays, seed=seed)
    elif kind == "colmap":
        o, d, scale = synthetic.colmap_rays(n_rays, seed=seed)
    elif kind == "dummy":   # tests/dummy/hotdog: 2 images of 200x200, camera_angle_x 0.6911112 (focal 277.78), radius 4.0311
        f = 0.5 * 200 / math.tan(0.5 * 0.6911112)
        a = synthetic.camera_rays(200, 200, f, [2.6, -1.9, 2.4])
        b = synthetic.camera_rays(200, 200, f, [-3.1, 1.2, 2.3])
        o, d = torch.cat([a[0], b[0]]), torch.cat([a[1], b[1]])
    else:
        raise ValueError(kind)
    g = torch.Generator().manual_seed(seed + 1)
    rgbs = torch.rand(o.size(0), 3, generator=g)
    return o, d, rgbs, scale


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md recipe).  Sampled in-process through NVML
    (nvidia_ml_py) every 50 ms; `nvidia-smi -lms` is the fallback.  The sampler is started BEFORE the warm-up and only samples
    taken after mark() are reported: a looping nvidia-smi was seen to block kernel launches for 80-100 ms, at its first query
    and now and then later, which a sub-second timed region cannot absorb; the NVML calls used here read two registers."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.stamps, self.proc, self.t_mark = gpu_index, [], [], None, 0.0
        self.nvml, self.stop_flag, self.t = None, threading.Event(), None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index()))
            pynvml.nvmlDeviceGetClockInfo(self.nvml[1], pynvml.NVML_CLOCK_SM)            # both queries must work, else fall back
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.nvml[1])
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self._physical_index())], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv, h = self.nvml
        bits = [(nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, mx, [name for bit, name in bits if mask & bit]))
                self.stamps.append(time.perf_counter())
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) > 8 and r[1].replace(".", "").isdigit():
                reasons = [nm for nm, v in zip(self.NAMES, r[5:9]) if v.lower().startswith("active")]
                mx = float(r[2]) if r[2].replace(".", "").isdigit() else None
                self.rows.append((float(r[1]), mx, reasons))
                self.stamps.append(time.perf_counter())

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML / nvidia-smi unavailable"]}
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
        if self.t is not None:
            self.t.join(timeout=2)
        keep = [r for r, ts in zip(self.rows, self.stamps) if ts >= self.t_mark]
        rows = keep if keep else self.rows[-1:]   # a region shorter than the sampling period: the closest sample
        sm = sorted(r[0] for r in rows)
        reasons = sorted({nm for r in rows for nm in r[2]})
        mx = max((r[1] for r in rows if r[1] is not None), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---- the reference's own formulation of a step (oracle/ref_port.py): CPU arm, cpu_baseline, gpu_reference ----
class ReferenceStep:
    """One training iteration in the reference's pure-PyTorch formulation (oracle/ref_port.py; weights op = the C
    restatement of src/cuda.cu on CPU, the UNMODIFIED reference kernel on CUDA) for every method / scene type of
    src/run.py:128-160.  Reported baseline, never shipped.  The occupancy grid is the workload's initial state and is not
    updated (a CPU update evaluates the field on 2.1 M points: far outside a bounded sample)."""

    def __init__(self, workload: str, target_samples: int, seed: int, device: str = "cpu"):
        from oracle import ref_port as rp
        from tinynerf_b200 import synthetic
        spec = TRAIN_SPECS[workload]
        self.rp, self.spec, self.dev = rp, spec, torch.device(device)
        dev = self.dev
        torch.manual_seed(seed)
        P = lambda t: torch.nn.Parameter(t.to(dev))
        lin = lambda i, o: (P(torch.randn(o, i) / math.sqrt(i)), P(torch.zeros(o)))
        mlp = lambda i, h, L, o=None: [lin(i, h)] + [lin(h, h) for _ in range(L)] + [lin(h, h if o is None else o)]
        m = spec["method"]
        if m == "kplanes":
            self.planes = [[P(torch.rand(1, 32, r, r)) for _ in range(3)] for r in (128, 256, 512)]
            feat, field_params = 96, [p for s in self.planes for p in s]
            self.feature_fn = lambda x: rp.kplanes_features(self.planes, x)
        elif m == "cobafa":
            res, ch = torch.linspace(32.0, 128, 6).int().tolist(), [8, 8, 8, 4, 4, 4]
            self.freqs = torch.linspace(2.0, 8.0, 6).tolist()
            self.basis = [P(torch.rand(1, c, r, r, r)) for r, c in zip(res, ch)]
            self.coef = P(torch.rand(1, 6, 64, 64, 64))
            self.trunk = mlp(36, 128, 5)
            feat, field_params = 128, self.basis + [self.coef] + [t for l in self.trunk for t in l]
            self.feature_fn = lambda x: rp.mlp(self.trunk, torch.nn.functional.dropout(
                rp.cobafa_lookup(self.basis, self.coef, self.freqs, x), 0.01, training=True))
        else:
            self.trunk = mlp(60, 256, 8)
            feat, field_params = 256, [t for l in self.trunk for t in l]
            self.feature_fn = lambda x: rp.mlp(self.trunk, rp.positional_encoding(x, 10))
        self.sig = mlp(feat, 64, 0, 1)
        self.col = mlp(feat + 51, 64, 3, 3)
        params = field_params + [t for l in self.sig + self.col for t in l]
        self.opt = torch.optim.Adam(params, lr=1e-2, eps=1e-15, weight_decay=1e-5)
        S = spec["n_samples"]
        self.S = S
        o, d, rgb, scale = make_scene(spec["rays"], 1 << 16, seed)
        self.o, self.d, self.rgb = o.to(dev), d.to(dev), rgb.to(dev)
        self.aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], device=dev)
        self.scene_scale = scale
        self.grid = (synthetic.analytic_grid(128, seed=SEED + 2) if spec["pin_grid"] else torch.ones(128, 128, 128)).to(dev)
        self.thr = min(0.01, self.grid.mean().item())
        self.pos, self.target = 0, target_samples
        self.chunk = max(64, min(BATCH, target_samples // 64))

    def step(self) -> int:
        rp, spec = self.rp, self.spec
        with torch.no_grad():  # dynamic batch accumulator, src/run.py:215-244
            cur, proj, k, ps, infos, rgbs = 0, 0, 0, [], [], []
            while proj < self.target:
                if self.pos + self.chunk > self.o.size(0):
                    self.pos = 0
                sl = slice(self.pos, self.pos + self.chunk)
                self.pos += self.chunk
                noise = torch.rand(self.chunk, self.S, device=self.dev)
                p, info, _ = rp.ray_provider(self.o[sl], self.d[sl], self.grid, self.thr, scene=spec["scene"], n_samples=self.S,
                                             aabb=self.aabb, near=0.1, far=1e5, uniform_range=self.scene_scale, noise=noise)
                info[:, 0] += cur
                ps.append(p); infos.append(info); rgbs.append(self.rgb[sl])
                cur += p.size(0); k += 1
                proj = int(cur * (1 + 1 / k))
            packed, info, rgb = torch.cat(ps), torch.cat(infos), torch.cat(rgbs)
        out = rp.render(self.feature_fn, lambda f: rp.sigma_head(self.sig, f), lambda f, dd: rp.rgb_head(self.col, 8, f, dd),
                        packed, info, torch.ones(3, device=self.dev))
        loss = torch.nn.functional.mse_loss(out, rgb)
        if spec["method"] == "kplanes":
            loss = loss + 1e-4 * rp.kplanes_tv(self.planes)
        self.opt.zero_grad()
        (loss * 1024.0).backward()   # GradScaler(2**10), never unscaled (src/run.py:201,259)
        self.opt.step()
        return packed.size(0)


def run_reference(workload: str, target_samples: int, steps: int, warmup: int, device: str = "cpu"):
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceStep(workload, target_samples, SEED, device)
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    for _ in range(warmup):
        ref.step()
    sync()
    t0, n = time.perf_counter(), 0
    for _ in range(steps):
        n += ref.step()
    sync()
    dt = time.perf_counter() - t0
    return n / dt, dt / steps * 1e3, n


def reference_render(n_chunks: int):
    """The reference's render loop (src/run.py:34-44) restated on the CPU: `n_chunks` chunks of 2048 rays of an 800x800 pose
    (K-Planes, analytic grid).  -> rays/s"""
    from oracle import ref_port as rp
    from tinynerf_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceStep("kplanes_aabb_2e18", 1 << 18, SEED, "cpu")
    o, d = synthetic.camera_rays(800, 800, 0.5 * 800 / math.tan(0.5 * 0.6911112), [2.6, -1.9, 2.4])
    mid = (o.size(0) // 2) // 2048 * 2048    # chunks from the middle rows of the image (they see the scene)
    t0 = time.perf_counter()
    with torch.no_grad():
        for k in range(mid, mid + n_chunks * 2048, 2048):
            packed, info, _ = rp.ray_provider(o[k:k + 2048], d[k:k + 2048], ref.grid, ref.thr, scene="aabb", n_samples=256,
                                              aabb=ref.aabb, near=0.1, far=1e5)
            if packed.size(0):
                rp.render(ref.feature_fn, lambda f: rp.sigma_head(ref.sig, f), lambda f, q: rp.rgb_head(ref.col, 8, f, q),
                          packed, info, torch.ones(3))
    dt = time.perf_counter() - t0
    return n_chunks * 2048 / dt, dt


# ---- weights microbench (config 5): device time of the launches, replayed from a CUDA graph ----------
def weights_microbench(dev, logn: int, peak: float, with_reference: bool = True):
    """fwd / bwd of the packed weights op at N = 2^logn.  The launches are captured in a CUDA graph and replayed, so the
    time is the kernels' (no Python between them); the graph cycles through enough independent input sets that the
    working set is > 2x the 126 MB L2 (every launch streams from HBM).  `reference_kernel`: the UNMODIFIED reference op
    (oracle/_ref/_cuda.so, thread-per-ray, incl. its zeros_like) on the same input sets, CUDA-event timed on its own
    (legacy) stream -- valid for R <= 2^20 rays (src/cuda.cu:81-86)."""
    from tinynerf_b200 import _cuda, synthetic
    n = 1 << logn
    n_sets = max(1, min(16, math.ceil(300e6 / (12 * n))))
    sets = []
    for i in range(n_sets):
        sig, info, g = synthetic.packed_rays(n, seed=1000 + logn + 7 * i)
        sig, info, g = sig.to(dev), info.to(dev), g.to(dev)
        sets.append((sig, torch.full_like(sig, 5.196 / 256), info, g))
    r = sets[0][2].size(0)
    out = {"n_samples": n, "n_rays": r, "input_sets": n_sets}
    reps = max(8, 2 * n_sets)
    for label, flags in (("", _cuda.TRUSTED_PARTITION), ("_unvalidated_info", 0)):
        ws = [_cuda.weights_fwd(sg, st, inf, 1e-4, flags) for sg, st, inf, _ in sets]
        for name, fn, nbytes in (("fwd", lambda i: _cuda.weights_fwd(sets[i][0], sets[i][1], sets[i][2], 1e-4, flags), 12 * n + 8 * r),
                                 ("bwd", lambda i: _cuda.weights_bwd(sets[i][0], sets[i][1], sets[i][2], ws[i], sets[i][3], flags), 20 * n + 8 * r)):
            fn(0)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for k in range(reps):
                    fn(k % n_sets)
            graph.replay()
            torch.cuda.synchronize()
            times = []
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); graph.replay(); e.record()
                torch.cuda.synchronize()
                times.append(s.elapsed_time(e) / reps)
            ms = sorted(times)[len(times) // 2]
            out[name + label] = {"us": round(ms * 1e3, 1), "GB/s": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 3)}
            del graph
    if with_reference and r <= (1 << 20):
        try:
            import oracle
            ref = oracle.load_ref_cuda()
            if ref is not None:
                ws = [ref.compute_weights_fwd(sg, st, inf, 1e-4) for sg, st, inf, _ in sets]
                rr = {}
                for name, fn, nbytes in (("fwd", lambda i: ref.compute_weights_fwd(sets[i][0], sets[i][1], sets[i][2], 1e-4), 12 * n + 8 * r),
                                         ("bwd", lambda i: ref.compute_weights_bwd(sets[i][0], sets[i][1], sets[i][2], ws[i], sets[i][3]), 20 * n + 8 * r)):
                    fn(0)
                    torch.cuda.synchronize()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    for k in range(reps):
                        fn(k % n_sets)
                    e.record()
                    torch.cuda.synchronize()
                    ms = s.elapsed_time(e) / reps
                    rr[name] = {"us": round(ms * 1e3, 1), "GB/s": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 3),
                                "speedup_ours": round(ms * 1e3 / out[name]["us"], 2)}
                out["reference_kernel"] = rr
        except Exception as ex:  # the checker is optional here
            out["reference_kernel"] = {"unavailable": repr(ex)[:200]}
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of one training
# iteration (profiles/r02_ncu_full.md, N = 234,218 packed samples): the `traffic` of the roofline object.
NCU_TRAFFIC_BYTES = {
    "tnf_kplanes_bwd": 643.3e6,         # kplanes_kernel<1,4>: 492.6 MB read + 150.7 MB written
    "tnf_kplanes_fwd": 255.9e6,         # kplanes_kernel<0,6>: 165.2 + 90.7 MB
    "tnf_heads_fwd": 444.3e6,           # 156.1 + 288.2 MB
    "tnf_heads_bwd_data": 606.6e6,      # 337.0 + 269.6 MB
    "tnf_linear_bwd_weight": 138.2e6,   # wgrad_tma_kernel, 64x64 layer: 134.5 + 3.7 MB
    "tnf_adam_step_grid": 869.7e6,      # 529.0 + 340.7 MB
    "tnf_tv_fwd_bwd": 215.7e6,          # tv_march_kernel: 139.1 + 76.6 MB
    "tnf_head_bwd": 96.7e6,             # head_bwd_kernel<3>: 73.5 + 23.2 MB
}
TENSOR_BOUND = {"tnf_heads_fwd", "tnf_heads_bwd_data", "tnf_linear_fwd", "tnf_linear_bwd_data", "tnf_wide_linear_fwd",
                "tnf_wide_linear_bwd_data", "tnf_wide_linear_bwd_weight"}


def summarise_profile(records, peak, tf32_peak):
    """records: (name, start, end, bytes, flops) -> per-entry-point totals and the dominant one."""
    agg = {}
    for name, s, e, nbytes, flops in records:
        a = agg.setdefault(name, {"ms": 0.0, "launches": 0, "bytes": 0, "flops": 0})
        a["ms"] += s.elapsed_time(e)
        a["launches"] += 1
        a["bytes"] += nbytes
        a["flops"] += flops
    table = {}
    for name, a in agg.items():
        gbs = a["bytes"] / a["ms"] / 1e6 if a["ms"] > 0 else 0.0
        row = {"launches": a["launches"], "avg_us": round(a["ms"] / a["launches"] * 1e3, 2),
               "total_ms": round(a["ms"], 3), "alg_MB_per_launch": round(a["bytes"] / a["launches"] / 1e6, 3),
               "GB/s": round(gbs, 1), "frac": round(gbs / peak, 4)}
        if a["flops"]:
            # issued tensor work = 3 TF32 MMAs per fp32-accurate product ("3xTF32")
            tf = a["flops"] / a["ms"] / 1e9
            row.update({"TFLOP/s_fp32_equiv": round(tf, 2), "TFLOP/s_tf32_issued": round(3 * tf, 2),
                        "tensor_frac_issued": round(3 * tf / tf32_peak, 4)})
        table[name] = row
    dom = max(table, key=lambda k: table[k]["total_ms"]) if table else None
    return table, dom


def roofline_of(table, dom, peak, tf32_peak, peak_src, samples_per_launch=0, ncu_applies=True):
    """ncu_applies: the committed ncu capture (NCU_TRAFFIC_BYTES) is of the K-Planes iteration; other workloads launch the same
    entry points at other shapes (Cobafa: 128-wide layers), so their `traffic` is left null rather than borrowed."""
    if not dom:
        return None
    t = table[dom]
    traffic = NCU_TRAFFIC_BYTES.get(dom) if ncu_applies else None
    if dom in TENSOR_BOUND and "TFLOP/s_tf32_issued" in t:
        return {"kernel": dom, "bound": "tensor", "achieved": t["TFLOP/s_tf32_issued"], "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": t["tensor_frac_issued"], "traffic": traffic, "peak_source": peak_src + " (bf16 burst / 2 = TF32)",
                "avg_us": t["avg_us"], "hbm_frac": t["frac"], "note": "3xTF32: three issued TF32 MMAs per fp32-accurate product"}
    roof = {"kernel": dom, "bound": "hbm", "achieved": t["GB/s"], "peak": peak, "unit": "GB/s", "frac": t["frac"],
            "traffic": traffic, "traffic_source": "profiles/r02_ncu_full.md" if traffic is not None else None,
            "peak_source": peak_src, "avg_us": t["avg_us"], "alg_bytes_per_launch": int(t["alg_MB_per_launch"] * 1e6)}
    if dom in ("tnf_kplanes_bwd", "tnf_kplanes_fwd") and samples_per_launch:
        # informational: the resource this kernel actually saturates.  Every sample moves 36 corner lines x 128 B through L2
        # each way; the L2 executes 128-byte-line reductions at 6.7 TB/s of payload and serves line gathers at 16 TB/s
        # (scripts/ubench/atomics_bench.cu, profiles/r02_atomics_ubench.txt) -- the HBM-byte accounting above cannot see that.
        payload = 36 * 128 * samples_per_launch
        l2_peak = 6700.0 if dom == "tnf_kplanes_bwd" else 16000.0
        roof["l2_line_traffic"] = {"what": "red.v4 payload (scatter half)" if dom == "tnf_kplanes_bwd" else "corner-line gathers",
                                   "bytes_per_launch": int(payload), "GB/s": round(payload / t["avg_us"] / 1e3, 1),
                                   "measured_l2_peak_GB/s": l2_peak, "frac": round(payload / t["avg_us"] / 1e3 / l2_peak, 3)}
    return roof


class Ctx:
    """Per-process state shared by the workloads."""

    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.rank, self.world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        # opt-in (TNF_BLOCKING_SYNC=1): measured at 8 ranks it trades the 30-60 ms scheduling stalls of the per-step-synchronised
        # arm (e2e 268 -> 468 M samples/s) for wake-up latency on every wait (value arm 746 -> 663 M samples/s)
        self.blocking_sync = set_blocking_sync(self.local_rank) if os.environ.get("TNF_BLOCKING_SYNC") == "1" else False
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peak, self.tf32_peak, self.peak_src = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, n: float, ms: float):
        """(sum over ranks of n, max over ranks of ms)"""
        if self.world == 1:
            return n, ms
        t = torch.tensor([float(n), ms], device=self.dev, dtype=torch.float64)
        a, b = t[0].clone(), t[1].clone()
        self.dist.all_reduce(a)
        self.dist.all_reduce(b, op=self.dist.ReduceOp.MAX)
        return float(a), float(b)


def train_workload(ctx: Ctx, name: str, steps: int, warmup: int, e2e_arm: bool, clocks=None):
    """One training workload: device-resident arm (value), a profiled pass (per-entry-point CUDA events -> kernels table,
    roofline) and, optionally, the end-to-end arm (host-resident rays, H2D per batch, loss read back)."""
    from tinynerf_b200 import _lib, synthetic
    from tinynerf_b200.run import RayStore, TrainConfig, Trainer
    spec = TRAIN_SPECS[name]
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    steps = max(steps, spec["min_steps"])
    n_store = 80_000 if spec["rays"] == "dummy" else N_STORE
    o, d, rgbs, scene_scale = make_scene(spec["rays"], n_store, SEED)
    cfg = TrainConfig(method=spec["method"], scene_type=spec["scene"], batch_size=BATCH, n_samples=spec["n_samples"],
                      scene_scale=scene_scale, seed=SEED, dp_mode=os.environ.get("TNF_DP_MODE", "peer"))
    if os.environ.get("TNF_OVERLAP_ADAM") == "1":   # diagnostics: A/B of the planes' Adam beside the weight-gradient kernels
        cfg.overlap_plane_adam = True
    analytic = synthetic.analytic_grid(128, seed=SEED + 2).to(dev) if spec["pin_grid"] else None
    analytic_mean = analytic.mean().item() if analytic is not None else None
    upd = {"events": [], "count": 0}

    def make_trainer(host: bool):
        torch.manual_seed(SEED)
        store = RayStore(o, d, rgbs, dev, host=host, seed=SEED, rank=rank, world=world)
        try:
            tr = Trainer(cfg, store, dev, rank=rank, world=world)
        except Exception as ex:   # symmetric memory unavailable on this box (no P2P / fabric): the NCCL baseline, loudly
            if world == 1 or cfg.dp_mode != "peer":
                raise
            sys.stderr.write(f"[bench] rank {rank}: peer-memory update unavailable ({ex!r}); falling back to dp_mode='nccl'\n")
            cfg.dp_mode = "nccl"
            DP_MODE_USED[0] = "nccl (peer-memory set-up failed)"
            tr = Trainer(cfg, store, dev, rank=rank, world=world)
        if analytic is not None:
            tr.occupancy_grid.grid.copy_(analytic)
            tr.occupancy_grid.mean = analytic_mean

            def pin_state(t):  # keep the occupancy state fixed (the update work itself has just been done inside step())
                t.occupancy_grid.grid.copy_(analytic)
                t.occupancy_grid.mean = analytic_mean
            tr.post_update = pin_state
        inner = tr.update_occupancy

        def timed_update():   # CUDA events around every occupancy update: its cost and how many fell into the window
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); inner(); e.record()
            upd["events"].append((s, e)); upd["count"] += 1
        tr.update_occupancy = timed_update
        return tr

    losses = []

    def one_step(tr, read_loss: bool):
        info = tr.step()
        if read_loss and tr.train_step >= 2:
            # D2H of the step's result: every iteration's loss is copied to pinned host memory (Trainer._publish_loss) and
            # consumed here one iteration late, so the host never drains the pipeline (the last one is read in timed())
            losses.append(tr.read_loss(tr.train_step - 2))
        return info["n_samples"]

    host_ms, host_dist = [], []

    def timed(tr, k, read_loss, profile):
        ctx.barrier()
        seg0 = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
        l0, u0 = _lib.launch_count, upd["count"]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            _lib.profile_start()
        s.record()
        n = 0
        h0 = time.perf_counter()
        per = []
        for _ in range(k):
            n += one_step(tr, read_loss)
            per.append(time.perf_counter())
        if read_loss:
            losses.append(tr.read_loss())   # the latest iteration's loss: inside the timed region
        if getattr(tr, "_fused", None) is not None:
            tr._fused.wait_updates()        # data-parallel: the last iteration's parameter updates (side stream) end inside the window
        e.record()
        host_ms.append((time.perf_counter() - h0) * 1e3 / k)  # host time per step (includes the batch-size sync)
        raw = [b - a for a, b in zip([h0] + per[:-1], per)]
        dd = sorted(raw)
        host_dist.append({"p50": round(dd[len(dd) // 2] * 1e3, 3), "p90": round(dd[int(len(dd) * 0.9)] * 1e3, 3),
                          "max": round(dd[-1] * 1e3, 3), "argmax": raw.index(dd[-1]),
                          "cudaMalloc_calls_in_window": int(torch.cuda.memory_stats(dev).get("segment.all.allocated", 0) - seg0)})
        ctx.barrier()
        recs = _lib.profile_stop() if profile else None
        ms = s.elapsed_time(e)
        upd_ms = [a.elapsed_time(b) for a, b in upd["events"][u0:]]
        n_all, ms_all = ctx.reduce(n, ms)
        return n_all, ms_all, recs, _lib.launch_count - l0, upd_ms

    # Before the W warm-up steps of each arm the trainer runs one full occupancy-update cycle untimed ("settle_steps"): the
    # update iteration allocates its temporaries while several steps are in flight, and the first time that happens the
    # caching allocator has to cudaMalloc (milliseconds, once per process) -- a short window would otherwise carry it.
    # Workloads whose grid evolves (not pinned) are timed from the start of training instead: their updates are the point.
    cadence = int(16 * 4096 / BATCH)
    settle = 0 if (os.environ.get("TNF_BENCH_NO_SETTLE") == "1" or not spec["pin_grid"]) else cadence + 2

    # ---- device-resident arm (value) ----
    tr = make_trainer(host=False)
    for _ in range(settle + warmup):
        one_step(tr, False)
    torch.cuda.synchronize()
    settle_upd_ms = [a.elapsed_time(b) for a, b in upd["events"]]
    if clocks is not None:
        clocks.mark()
    n, ms, _, launches, upd_ms = timed(tr, steps, False, profile=False)
    clk = clocks.stop() if clocks is not None else None
    value = n / (ms * 1e-3)
    occupancy_now = tr.occupancy_grid.occupancy()
    # same trainer, same timed-region structure, now with a CUDA-event pair around every C-ABI call on its
    # launching stream (kept out of the headline pass: ~30 extra event records per step)
    _, _, recs, _, _ = timed(tr, min(steps, 16), False, profile=True)
    table, dom = summarise_profile(recs, ctx.peak, ctx.tf32_peak)
    tr.close()
    del tr
    torch.cuda.empty_cache()
    res = {"workload": name, "value": round(value, 1), "unit": UNIT, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms / steps, 4), "packed_samples_per_step_per_gpu": round(n / steps / world),
           "gpu_launches": int(launches), "host_ms_per_step": round(host_ms[0], 4), "host_step_ms": {"value_arm": host_dist[0]},
           "settle_steps": settle, "clocks": clk, "kernels": table,
           "roofline": roofline_of(table, dom, ctx.peak, ctx.tf32_peak, ctx.peak_src, n / steps / world,
                                   ncu_applies=spec["method"] == "kplanes")}
    # occupancy update inside / outside the window (SURVEY 8d defines the metric with the update amortised at its cadence)
    all_upd = upd_ms or settle_upd_ms[-1:]
    upd_avg = sum(all_upd) / len(all_upd) if all_upd else None
    occ = {"cadence_steps": cadence, "updates_in_timed_window": len(upd_ms), "occupied_frac_after": round(occupancy_now, 4),
           "update_ms": None if upd_avg is None else round(upd_avg, 3),
           "update_ms_source": "timed window" if upd_ms else "the settle cycle before the window"}
    if upd_avg is not None:
        base = (ms - sum(upd_ms)) / steps          # ms per step without any update
        occ["value_amortised_at_cadence"] = round(n / steps / ((base + upd_avg / cadence) * 1e-3), 1)
    res["occupancy_update"] = occ

    # ---- end-to-end arm (host buffers, H2D per batch, loss read back) ----
    if e2e_arm:
        tr = make_trainer(host=True)
        for _ in range(settle + warmup):
            one_step(tr, True)
        h0 = tr.store.h2d_bytes
        n2, ms2, _, _, _ = timed(tr, steps, True, profile=False)
        e2e = n2 / (ms2 * 1e-3)
        h2d = (tr.store.h2d_bytes - h0) / steps
        batches_per_step = h2d / (BATCH * 44)   # 36 B of ray data + an 8-byte index per ray
        d2h = 4 + 8 * batches_per_step  # loss + one packed-sample count per provider call
        tr.close()
        del tr
        torch.cuda.empty_cache()
        res["e2e"] = {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": round(ms2 / steps, 4)}
        res["host_step_ms"]["e2e_arm"] = host_dist[-1]
    return res


def weights_workload(ctx: Ctx, logns=(18, 22, 24, 26)):
    """Config 5 at every N: each rank runs the microbench on its own GPU with its own inputs; aggregate GB/s = sum over ranks
    of bytes / max over ranks of time, per size and direction."""
    micro = {f"2^{ln}": weights_microbench(ctx.dev, ln, ctx.peak, with_reference=ctx.rank == 0) for ln in logns}
    agg = {}
    for key, m in micro.items():
        row = {}
        for d_, nb in (("fwd", 12 * m["n_samples"] + 8 * m["n_rays"]), ("bwd", 20 * m["n_samples"] + 8 * m["n_rays"])):
            n_all, us_all = ctx.reduce(nb, m[d_]["us"])
            row[d_] = {"aggregate_GB/s": round(n_all / us_all / 1e3, 1), "frac_of_n_gpus_x_peak": round(n_all / us_all / 1e3 / (ctx.peak * ctx.world), 3)}
        agg[key] = row
    big = micro[f"2^{logns[-1]}"]
    nb = 32 * big["n_samples"] + 16 * big["n_rays"]
    n_all, us_all = ctx.reduce(nb, big["fwd"]["us"] + big["bwd"]["us"])
    return {"workload": "weights_micro", "value": round(n_all / us_all / 1e3, 1), "unit": "GB/s (fwd+bwd, largest size, all ranks)",
            "frac_of_peak": round(n_all / us_all / 1e3 / (ctx.peak * ctx.world), 3), "aggregate": agg, "rank0": micro}


def render_workload(ctx: Ctx, reps: int = 5):
    """The render half (src/run.py:15-50): whole 800x800 poses through Trainer.render (one host sync per image, fixed-capacity
    sample buffer, forward-only kernels); rays uploaded from the host and the image read back inside the timed region."""
    from tinynerf_b200 import _lib, synthetic
    from tinynerf_b200.run import RayStore, TrainConfig, Trainer
    dev = ctx.dev
    o, d, rgbs, _ = make_scene("blender", 1 << 14, SEED)
    torch.manual_seed(SEED)
    tr = Trainer(TrainConfig(method="kplanes", scene_type="aabb", batch_size=BATCH, n_samples=N_SAMPLES, seed=SEED),
                 RayStore(o, d, rgbs, dev, seed=SEED), dev)
    analytic = synthetic.analytic_grid(128, seed=SEED + 2).to(dev)
    tr.occupancy_grid.grid.copy_(analytic)
    tr.occupancy_grid.mean = analytic.mean().item()
    focal = 0.5 * 800 / math.tan(0.5 * 0.6911112)
    poses = [synthetic.camera_rays(800, 800, focal, [4.0311 * math.cos(a) * 0.8, 4.0311 * math.sin(a) * 0.8, 4.0311 * 0.6])
             for a in (0.3, 1.4, 2.9, 4.1, 5.2)]
    poses = [(a.pin_memory(), b.pin_memory()) for a, b in poses]
    host_img = torch.empty(640000, 3).pin_memory()
    n_samples = []

    def image(i):
        ro, rd = poses[i % len(poses)]
        img = tr.render(ro, rd)
        host_img.copy_(img, non_blocking=True)

    image(0); image(1)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        image(i)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    launches = (_lib.launch_count - l0) / reps
    _lib.profile_start()
    image(0)
    torch.cuda.synchronize()
    recs = _lib.profile_stop()
    table, dom = summarise_profile(recs, ctx.peak, ctx.tf32_peak)
    n_samp = sum(r[3] for r in recs if r[0] == "tnf_composite_fwd") // 16   # composite's algorithmic bytes ~ 16 B/sample
    tr.close()
    return {"workload": "render_800", "value": round(640000 / (ms * 1e-3), 1), "unit": "rays/s", "ms_per_image": round(ms, 3),
            "packed_samples_per_image": int(n_samp), "samples_per_s": round(n_samp / (ms * 1e-3), 1),
            "h2d_bytes_per_image": 2 * 640000 * 12, "d2h_bytes_per_image": 640000 * 12 + 8 * 313, "host_syncs_per_image": 1,
            "gpu_launches_per_image": round(launches, 1), "kernels": table,
            "roofline": roofline_of(table, dom, ctx.peak, ctx.tf32_peak, ctx.peak_src)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kplanes_aabb_2e18", choices=WORKLOADS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other workloads, gpu_reference and parity")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    wl = args.workload
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        if wl == "render_800":
            v, dt = reference_render(8)
            line = {"impl": "reference", "metric": "rays/sec rendered", "value": round(v, 1), "unit": "rays/s", "n_gpus": args.gpus,
                    "steps": 1, "warmup": 0, "ms_per_step": round(dt * 1e3, 1), "config": {"workload": wl},
                    "cpu_baseline": {"value": round(v, 1), "unit": "rays/s", "cores": cores, "kind": "port",
                                     "sample": "8 chunks of 2048 rays from the middle rows of one 800x800 pose"}}
        elif wl == "weights_micro":
            from oracle import c as orc
            from tinynerf_b200 import synthetic
            sig, info, g = synthetic.packed_rays(1 << 24, seed=1024)
            steps_t = torch.full_like(sig, 5.196 / 256)
            t0 = time.perf_counter()
            w = orc.weights_fwd(sig, steps_t, info, 1e-4)
            orc.weights_bwd(sig, steps_t, info, w, g)
            dt = time.perf_counter() - t0
            v = (32 * sig.numel() + 16 * info.size(0)) / dt / 1e9
            line = {"impl": "reference", "metric": "weights-kernel HBM GB/s", "value": round(v, 3), "unit": "GB/s", "n_gpus": args.gpus,
                    "steps": 1, "warmup": 0, "ms_per_step": round(dt * 1e3, 1), "config": {"workload": wl},
                    "cpu_baseline": {"value": round(v, 3), "unit": "GB/s", "cores": 1, "kind": "port",
                                     "sample": "fwd+bwd of 2^24 packed samples, C restatement of src/cuda.cu (scalar, 1 thread)"}}
        else:
            tgt = BATCH * TRAIN_SPECS[wl]["n_samples"]
            v, ms, n = run_reference(wl, tgt, args.steps, args.warmup)
            sample = (f"{args.steps} full steps of ~{tgt} packed samples each (the same per-step target as the GPU arm), "
                      f"occupancy grid at the workload's initial state, no update inside; torch threads={cores}")
            line = {"impl": "reference", "metric": METRIC, "value": round(v, 1), "unit": UNIT, "n_gpus": args.gpus,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
                    "config": {"workload": wl, "per_step_samples": tgt, "rays_per_chunk": BATCH,
                               "samples_per_ray": TRAIN_SPECS[wl]["n_samples"]},
                    "cpu_baseline": {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}}
        line.update({"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                     "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return

    ctx = Ctx(args)
    dist = ctx.dist
    clocks = ClockSampler(ctx.local_rank) if rank == 0 else None
    if clocks is not None:
        clocks.start()

    def finish(line):
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()

    base = {"n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}

    if wl == "weights_micro":
        r = weights_workload(ctx)
        clk = clocks.stop() if clocks is not None else None
        finish({"metric": "weights-kernel HBM GB/s", "value": r["value"], "unit": "GB/s", **base, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": None, "config": {"workload": wl, "sizes": list(r["rank0"]), "l2": "rotating input sets > 2x L2",
                                                 "parallelism": f"independent inputs per rank x{world}"},
                "clocks": clk, "roofline": {"kernel": "tnf_weights_fwd+bwd", "bound": "hbm", "achieved": r["value"], "peak": ctx.peak * world,
                                            "unit": "GB/s", "frac": r["frac_of_peak"], "traffic": None, "peak_source": ctx.peak_src},
                "weights_microbench": r})
        return
    if wl == "render_800":
        r = render_workload(ctx)
        clk = clocks.stop() if clocks is not None else None
        finish({"metric": "rays/sec rendered", "value": r["value"], "unit": "rays/s", **base, "n_gpus": 1, "steps": 5, "warmup": 2,
                "ms_per_step": r["ms_per_image"], "config": {"workload": wl, "image": "800x800", "chunking": "one count launch per image, <= 2^20-sample chunks"},
                "clocks": clk, "roofline": r["roofline"], "render": r,
                "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": r["h2d_bytes_per_image"], "d2h_bytes_per_step": r["d2h_bytes_per_image"]}})
        return

    # ---- a training workload is the headline ----
    main_res = train_workload(ctx, wl, args.steps, args.warmup, e2e_arm=True, clocks=clocks)
    spec = TRAIN_SPECS[wl]
    extras, micro, cpu, gpu_ref, par = {}, None, None, None, None
    if not args.no_extras:
        # short runs of the other BASELINE configurations at this N (value arm only), each isolated: a failure is recorded,
        # the headline line is printed regardless
        for other, k in (("cobafa_aabb_dyn", 24), ("kplanes_unbounded_decay", 200), ("kplanes_aabb_2e18", 64)):
            if other == wl:
                continue
            try:
                r = train_workload(ctx, other, k, 3, e2e_arm=False)
                extras[other] = {kk: r[kk] for kk in ("value", "unit", "steps", "ms_per_step", "packed_samples_per_step_per_gpu",
                                                      "gpu_launches", "occupancy_update", "roofline", "kernels")}
            except Exception as ex:
                extras[other] = {"error": repr(ex)[:300], "trace": traceback.format_exc()[-600:]}
                ctx.barrier()
    if not args.no_microbench:
        try:
            micro = weights_workload(ctx)
        except Exception as ex:
            micro = {"error": repr(ex)[:300]}
    if rank == 0 and world == 1 and not args.no_extras:
        for nm, fn in (("render_800", lambda: render_workload(ctx)), ("vanilla_dummy", lambda: train_workload(ctx, "vanilla_dummy", 16, 3, e2e_arm=False))):
            if nm == wl:
                continue
            try:
                r = fn()
                extras[nm] = {kk: r[kk] for kk in r if kk not in ("clocks", "host_step_ms", "warmup", "settle_steps")}
            except Exception as ex:
                extras[nm] = {"error": repr(ex)[:300], "trace": traceback.format_exc()[-600:]}
        try:   # informational: the reference's formulation on this GPU (stock torch ops + the reference weights kernel)
            v, gms, _ = run_reference(wl, BATCH * spec["n_samples"], 5, 2, device="cuda")
            gpu_ref = {"value": round(v, 1), "unit": UNIT, "ms_per_step": round(gms, 2), "kind": "port on cuda",
                       "what": "oracle/ref_port.py step on the B200: grid_sample / Linear / index_add_ / torch.optim.Adam + oracle/_ref/_cuda.so"}
        except Exception as ex:
            gpu_ref = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
        try:   # achieved worst-case errors of the CUDA path against the restatement (the checker; tests assert the bars)
            from oracle import parity
            par = {"march_pack": parity.march_parity(), "weights_trusted_partition_2^20": parity.weights_parity()}
            fsp = parity.fused_step_parity()
            fsp.pop("per_parameter", None)
            par["fused_step_2^18"] = fsp
            par["bars"] = "masks/packing bit-exact; weights, colours, loss 1e-5 rel; gradients rel-L2 2e-5, worst 5e-5 of tensor max (DESIGN section 2)"
        except Exception as ex:
            par = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tgt = BATCH * spec["n_samples"]
        v, cms, cn = run_reference(wl, tgt, 2, 1)
        cpu = {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 full steps (~{tgt} packed samples each) after 1 warm-up, {cms:.0f} ms/step, torch threads={cores}"}

    line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, **base, "steps": main_res["steps"], "warmup": args.warmup,
            "ms_per_step": main_res["ms_per_step"],
            "config": {"workload": wl, "rays_per_chunk": BATCH, "samples_per_ray": spec["n_samples"],
                       "packed_samples_per_step_per_gpu": main_res["packed_samples_per_step_per_gpu"],
                       "grid": "128^3 analytic ball+torus, pinned after every update" if spec["pin_grid"] else "128^3, starts all-ones, evolves (updates + decay)",
                       "l2": "inputs change every step (fresh rays; parameters + gradients + Adam state stream through L2 > 126 MB)",
                       "parallelism": f"ray-sharded dp{world}" + ("" if world == 1 else ", parameter update: " + (
                           "one reduce+Adam+broadcast kernel over NVLink peer memory" if DP_MODE_USED[0] == "peer"
                           else "NCCL all-reduce + replicated Adam [" + DP_MODE_USED[0] + "]")), "host_wait": "blocking" if ctx.blocking_sync else "spin",
                       "settle_steps": main_res["settle_steps"],
                       "e2e_loss_readback": "every step, async D2H into a pinned ring, consumed one step late"},
            "e2e": main_res.get("e2e"), "gpu_launches": main_res["gpu_launches"], "host_ms_per_step": main_res["host_ms_per_step"],
            "host_step_ms": main_res["host_step_ms"], "clocks": main_res["clocks"], "roofline": main_res["roofline"],
            "occupancy_update": main_res["occupancy_update"], "kernels": main_res["kernels"], "weights_microbench": micro,
            "cpu_baseline": cpu, "gpu_reference": gpu_ref, "parity": par, "workloads": extras}
    finish(line)


if __name__ == "__main__":
    main()
