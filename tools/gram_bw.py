"""Dense Gram build bandwidth (HBM-write bound): python tools/gram_bw.py  -> JSON lines.
DCGP_GRAM_RT=1 reproduces the one-row-tile-per-CTA grid."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops  # noqa: E402
from deconditional_downscaling_b200.exact import rbf_spec  # noqa: E402

dev = torch.device("cuda", 0)
for dt, N, d in [(torch.float32, 8192, 5), (torch.float32, 32768, 3), (torch.float64, 16384, 3), (torch.float64, 32768, 3)]:
    x = torch.randn(N, d, dtype=dt, device=dev)
    spec = rbf_spec(d, dt, dev)
    K = torch.empty(N, N, dtype=dt, device=dev)
    best = 1e9
    for i in range(6):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.gram(spec, x, None, diag_const=0.5, out=K)
        e.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, s.elapsed_time(e) * 1e-3)
    size = K.element_size()
    print(json.dumps({"kernel": "gram_fwd", "dtype": str(dt).split(".")[-1], "N": N, "d": d, "ms": best * 1e3,
                      "GBps": N * N * size / best / 1e9, "rt": os.environ.get("DCGP_GRAM_RT", "default(8)"),
                      "checksum": float(K[::997, ::991].double().sum().item())}), flush=True)
    del K
