"""Print the headline + per-kernel table of a bench.py log: python scripts/show_bench.py gpurun_out/benchNN.log"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith('{"metric'):
            continue
        d = json.loads(line)
        print(f"{path}: value {d['value']/1e6:.1f} M/s  {d['ms_per_step']} ms/step | e2e {d['e2e']['value']/1e6:.1f} M/s {d['e2e']['ms_per_step']} ms | "
              f"launches {d['gpu_launches']} | clocks {d['clocks']}")
        if d.get("roofline"):
            print("  roofline", d["roofline"])
        if d.get("weights_microbench"):
            print("  micro", d["weights_microbench"])
        if d.get("cpu_baseline"):
            print("  cpu", d["cpu_baseline"])
        tot = 0.0
        steps = min(d["steps"], 16)
        for k, v in (d.get("kernels") or {}).items():
            ms = v["total_ms"] / steps
            tot += ms
            extra = f" tensor {v['tensor_frac_issued']}" if "tensor_frac_issued" in v else ""
            print(f"  {k:24s} {v['launches']:4d} x {v['avg_us']:8.1f} us = {ms:7.3f} ms/step  hbm frac {v['frac']}{extra}")
        print(f"  sum of kernels {tot:.3f} ms/step")
        for name, w in (d.get("workloads") or {}).items():
            if "error" in w:
                print(f"  workload {name}: ERROR {w['error']}")
                continue
            rf = w.get("roofline") or {}
            print(f"  workload {name}: {w['value']/1e6:.2f} M {w['unit']}  {w.get('ms_per_step', w.get('ms_per_image'))} ms | dominant {rf.get('kernel')} "
                  f"{rf.get('bound')} frac {rf.get('frac')} | updates {w.get('occupancy_update')}")
        for key in ("gpu_reference", "parity", "occupancy_update"):
            if d.get(key):
                print(f"  {key}", d[key])
