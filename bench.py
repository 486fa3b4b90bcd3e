#!/usr/bin/env python
"""bench.py -- ray samples/sec per train step of the packed-ray hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload = "kplanes_aabb_2e18"): BASELINE config #2 -- K-Planes + vanilla heads on a
synthetic 800x800 Blender-shaped scene, AABB +-1.5, 128^3 occupancy grid held at a seeded analytic state
(ball + torus; the cadence update IS executed, then the state is restored so every step sees the same
occupancy), dynamic batches of 1024-ray chunks x 256 samples -> ~2^18 packed samples per step per GPU.
One "step" = march+pack -> K-Planes gather -> heads -> weights -> composite -> MSE+TV -> backward -> Adam
(+ the occupancy update when the reference's cadence says so).  Rays shard across ranks (weak scaling).

value       : packed samples processed by all ranks / max-over-ranks device time, ray store in HBM
e2e         : same through the public API with the ray store in pinned HOST memory (per-batch host
              gather + H2D inside the timed region) and the loss read back (D2H) every step
roofline    : the dominant kernel of ours in the step, CUDA-event timed per launch in the timed region
cpu_baseline: the oracle's PyTorch-CPU restatement of the same step on this box's host cores (rank 0)
--impl reference: that CPU path alone, as the reference arm.
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
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ray samples/sec per train step"
UNIT = "samples/s"
N_STORE = 1 << 21          # rays in the synthetic scene store (x36 B = 75 MB; > one epoch of the run)
BATCH, N_SAMPLES = 1024, 256
SEED = 1234


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


# ---- synthetic scene -------------------------------------------------------
def make_scene(n_rays: int, seed: int):
    from tinynerf_b200 import synthetic
    o, d = synthetic.blender_rays(n_rays, seed=seed)
    # colours of the analytic scene: white background, reddish ball / bluish torus hit test along the ray
    g = torch.Generator().manual_seed(seed + 1)
    rgbs = torch.rand(n_rays, 3, generator=g)
    return o, d, rgbs


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


# ---- the reference arm / cpu_baseline: the oracle's CPU restatement of the same step -------------
class CpuReferenceStep:
    """K-Planes AABB training iteration in the reference's own pure-PyTorch formulation on the host cores
    (oracle/ref_port.py, weights op = C restatement of src/cuda.cu).  Reported baseline, never shipped."""

    def __init__(self, target_samples: int, seed: int):
        from oracle import ref_port as rp
        from tinynerf_b200 import synthetic
        self.rp = rp
        torch.manual_seed(seed)
        self.aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
        self.planes = [[torch.nn.Parameter(torch.rand(1, 32, r, r)) for _ in range(3)] for r in (128, 256, 512)]
        lin = lambda i, o: (torch.nn.Parameter(torch.randn(o, i) / math.sqrt(i)), torch.nn.Parameter(torch.zeros(o)))
        self.sig = [lin(96, 64), lin(64, 1)]
        self.col = [lin(147, 64), lin(64, 64), lin(64, 64), lin(64, 64), lin(64, 3)]
        params = [p for s in self.planes for p in s] + [t for l in self.sig + self.col for t in l]
        self.opt = torch.optim.Adam(params, lr=1e-2, eps=1e-15, weight_decay=1e-5)
        self.grid = synthetic.analytic_grid(128, seed=SEED + 2)
        self.thr = min(0.01, self.grid.mean().item())
        self.o, self.d, self.rgb = make_scene(1 << 16, seed)
        self.pos, self.target = 0, target_samples
        self.chunk = max(64, min(BATCH, target_samples // 64))

    def step(self) -> int:
        rp = self.rp
        with torch.no_grad():  # dynamic batch accumulator, src/run.py:215-244
            cur, proj, k, ps, infos, rgbs = 0, 0, 0, [], [], []
            while proj < self.target:
                if self.pos + self.chunk > self.o.size(0):
                    self.pos = 0
                sl = slice(self.pos, self.pos + self.chunk)
                self.pos += self.chunk
                noise = torch.rand(self.chunk, N_SAMPLES)
                p, info, _ = rp.ray_provider(self.o[sl], self.d[sl], self.grid, self.thr, scene="aabb",
                                             n_samples=N_SAMPLES, aabb=self.aabb, near=0.1, far=1e5, noise=noise)
                info[:, 0] += cur
                ps.append(p); infos.append(info); rgbs.append(self.rgb[sl])
                cur += p.size(0); k += 1
                proj = int(cur * (1 + 1 / k))
            packed, info, rgb = torch.cat(ps), torch.cat(infos), torch.cat(rgbs)
        out = rp.render(lambda x: rp.kplanes_features(self.planes, x), lambda f: rp.sigma_head(self.sig, f),
                        lambda f, dd: rp.rgb_head(self.col, 8, f, dd), packed, info, torch.ones(3))
        loss = torch.nn.functional.mse_loss(out, rgb) + 1e-4 * rp.kplanes_tv(self.planes)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return packed.size(0)


def run_cpu(target_samples: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    ref = CpuReferenceStep(target_samples, SEED)
    for _ in range(warmup):
        ref.step()
    t0, n = time.perf_counter(), 0
    for _ in range(steps):
        n += ref.step()
    dt = time.perf_counter() - t0
    return n / dt, dt / steps * 1e3, n


# ---- weights microbench (config 5): device time of the launches, replayed from a CUDA graph ----------
def weights_microbench(dev, logn: int, peak: float):
    """fwd / bwd of the packed weights op at N = 2^logn.  The launches are captured in a CUDA graph and replayed, so the
    time is the kernels' (no Python between them); the graph cycles through enough independent input sets that the
    working set is > 2x the 126 MB L2 (every launch streams from HBM)."""
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
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of one training
# iteration (profiles/r01_ncu_full.md, N = 233,625 packed samples): the `traffic` of the roofline object.
NCU_TRAFFIC_BYTES = {
    "tnf_kplanes_bwd": 639.1e6,         # kplanes_kernel<1>: 488.1 MB read + 151.0 MB written
    "tnf_kplanes_fwd": 230.9e6,         # kplanes_kernel<0>: 151.8 + 79.1 MB
    "tnf_heads_fwd": 390.7e6,           # 138.9 + 251.8 MB
    "tnf_heads_bwd_data": 596.6e6,      # 331.3 + 265.3 MB
    "tnf_linear_bwd_weight": 136.5e6,   # wgrad_tma_kernel, 64x64 layer: 132.0 + 4.5 MB
    "tnf_adam_step_grid": 871.0e6,           # 529.1 + 341.9 MB
    "tnf_tv_fwd_bwd": 215.5e6,          # tv_march_kernel: 138.7 + 76.8 MB
    "tnf_head_bwd": 94.0e6,             # head_bwd_kernel<3>: 72.2 + 21.8 MB
}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return
        tgt = 1 << 16
        v, ms, n = run_cpu(tgt, args.steps, args.warmup)
        cores = os.cpu_count() or 1
        sample = f"{args.steps} steps of ~2^16 packed samples each (1/4 of the 2^18 workload per step), torch threads={cores}"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(v, 1), "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": "kplanes_aabb_2e18", "per_step_samples": tgt},
                          "cpu_baseline": {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": round(v, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    from tinynerf_b200 import _lib, synthetic
    from tinynerf_b200.run import RayStore, TrainConfig, Trainer
    # opt-in (TNF_BLOCKING_SYNC=1): measured at 8 ranks it trades the 30-60 ms scheduling stalls of the per-step-synchronised
    # arm (e2e 268 -> 468 M samples/s) for wake-up latency on every wait (value arm 746 -> 663 M samples/s)
    blocking_sync = set_blocking_sync(local_rank) if os.environ.get("TNF_BLOCKING_SYNC") == "1" else False
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, tf32_peak, peak_src = measured_peaks()

    o, d, rgbs = make_scene(N_STORE, SEED)
    analytic = synthetic.analytic_grid(128, seed=SEED + 2).to(dev)
    analytic_mean = analytic.mean().item()
    cfg = TrainConfig(method="kplanes", scene_type="aabb", batch_size=BATCH, n_samples=N_SAMPLES, seed=SEED)
    if os.environ.get("TNF_OVERLAP_ADAM") == "1":   # diagnostics: A/B of the planes' Adam beside the weight-gradient kernels
        cfg.overlap_plane_adam = True

    def make_trainer(host: bool):
        torch.manual_seed(SEED)
        store = RayStore(o, d, rgbs, dev, host=host, seed=SEED, rank=rank, world=world)
        tr = Trainer(cfg, store, dev, rank=rank, world=world)
        tr.occupancy_grid.grid.copy_(analytic)
        tr.occupancy_grid.mean = analytic_mean

        def pin_state(t):  # keep the occupancy state fixed (the update work itself has just been done inside step())
            t.occupancy_grid.grid.copy_(analytic)
            t.occupancy_grid.mean = analytic_mean
        tr.post_update = pin_state
        return tr

    def one_step(tr, read_loss: bool):
        info = tr.step()
        if read_loss and tr.train_step >= 2:
            # D2H of the step's result: every iteration's loss is copied to pinned host memory (Trainer._publish_loss) and
            # consumed here one iteration late, so the host never drains the pipeline (the last one is read in timed())
            losses.append(tr.read_loss(tr.train_step - 2))
        return info["n_samples"]

    losses = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms, host_dist = [], []

    def timed(tr, steps, read_loss, profile):
        barrier()
        l0 = _lib.launch_count
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            _lib.profile_start()
        s.record()
        n = 0
        h0 = time.perf_counter()
        per = []
        for _ in range(steps):
            n += one_step(tr, read_loss)
            per.append(time.perf_counter())
        if read_loss:
            losses.append(tr.read_loss())   # the latest iteration's loss: inside the timed region
        e.record()
        host_ms.append((time.perf_counter() - h0) * 1e3 / steps)  # host time per step (includes the batch-size sync)
        raw = [b - a for a, b in zip([h0] + per[:-1], per)]
        d = sorted(raw)
        host_dist.append({"p50": round(d[len(d) // 2] * 1e3, 3), "p90": round(d[int(len(d) * 0.9)] * 1e3, 3),
                          "max": round(d[-1] * 1e3, 3), "argmax": raw.index(d[-1])})
        barrier()
        recs = _lib.profile_stop() if profile else None
        ms = s.elapsed_time(e)
        tot = torch.tensor([float(n), ms], device=dev, dtype=torch.float64)
        if world > 1:
            n_all = tot[0].clone(); dist.all_reduce(n_all)
            ms_all = tot[1].clone(); dist.all_reduce(ms_all, op=dist.ReduceOp.MAX)
            n, ms = float(n_all), float(ms_all)
        return n, ms, recs, _lib.launch_count - l0

    # Before the W warm-up steps of each arm the trainer runs one full occupancy-update cycle untimed ("settle_steps"): the
    # update iteration allocates its temporaries while several steps are in flight, and the first time that happens the
    # caching allocator has to cudaMalloc (milliseconds, once per process) -- a 64-step window would otherwise carry it.
    settle = 0 if os.environ.get("TNF_BENCH_NO_SETTLE") == "1" else (1 << 16) // BATCH + 2

    # ---- device-resident arm (value) ----
    tr = make_trainer(host=False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(settle + args.warmup):
        one_step(tr, False)
    clocks.mark()
    n, ms, _, launches = timed(tr, args.steps, False, profile=False)
    clk = clocks.stop() if rank == 0 else None
    value = n / (ms * 1e-3)
    # same trainer, same timed-region structure, now with a CUDA-event pair around every C-ABI call on its
    # launching stream (kept out of the headline pass: ~30 extra event records per step)
    _, _, recs, _ = timed(tr, min(args.steps, 16), False, profile=True)
    table, dom = summarise_profile(recs, peak, tf32_peak)
    del tr
    torch.cuda.empty_cache()

    # ---- end-to-end arm (host buffers, H2D per batch, loss read back) ----
    tr = make_trainer(host=True)
    for _ in range(settle + args.warmup):
        one_step(tr, True)
    h0 = tr.store.h2d_bytes
    n2, ms2, _, _ = timed(tr, args.steps, True, profile=False)
    e2e = n2 / (ms2 * 1e-3)
    h2d = (tr.store.h2d_bytes - h0) / args.steps
    batches_per_step = h2d / (BATCH * 36)
    d2h = 4 + 8 * batches_per_step  # loss + one packed-sample count per provider call
    del tr
    torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    micro = None if args.no_microbench else {f"2^{ln}": weights_microbench(dev, ln, peak) for ln in (18, 22, 24, 26)}
    cpu = None
    if not args.no_cpu_baseline:
        v, cms, cn = run_cpu(1 << 18, 2, 1)
        cpu = {"value": round(v, 1), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"2 full steps (~2^18 packed samples each) after 1 warm-up, {cms:.0f} ms/step, torch threads={os.cpu_count()}"}

    roof = None
    if dom:
        t = table[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": t["GB/s"], "peak": peak, "unit": "GB/s", "frac": t["frac"],
                "traffic": NCU_TRAFFIC_BYTES.get(dom), "traffic_source": "profiles/r01_ncu_full.md" if dom in NCU_TRAFFIC_BYTES else None,
                "peak_source": peak_src, "avg_us": t["avg_us"],
                "alg_bytes_per_launch": int(t["alg_MB_per_launch"] * 1e6)}
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "kplanes_aabb_2e18", "rays_per_chunk": BATCH, "samples_per_ray": N_SAMPLES,
                       "packed_samples_per_step_per_gpu": round(n / args.steps / world), "grid": "128^3 analytic ball+torus",
                       "l2": "inputs change every step (fresh rays; 396 MB of plane params+grads+Adam state stream through L2 > 126 MB)",
                       "parallelism": f"ray-sharded dp{world}", "host_wait": "blocking" if blocking_sync else "spin",
                       "settle_steps": settle,
                       "e2e_loss_readback": "every step, async D2H into a pinned ring, consumed one step late"},
            "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": round(ms2 / args.steps, 4)},
            "gpu_launches": int(launches), "host_ms_per_step": round(host_ms[0], 4), "host_step_ms": {"value_arm": host_dist[0], "e2e_arm": host_dist[-1]}, "clocks": clk, "roofline": roof, "kernels": table,
            "weights_microbench": micro, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
