"""Diagnostics: the packed weights op alone (config 5) at 2^18..2^26 -- bench.weights_microbench (CUDA-graph replay over
rotating input sets larger than L2), trusted-partition and reference-signature (validated) paths, fractions of the
measured HBM peak.  `python scripts/time_weights.py [18 20 22 24 26]`"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench

peak = bench.measured_peaks()[0]
logns = [int(a) for a in sys.argv[1:]] or [18, 20, 22, 24, 26]
for ln in logns:
    r = bench.weights_microbench(torch.device("cuda"), ln, peak, with_reference=False)
    print(json.dumps({f"2^{ln}": r}), flush=True)
