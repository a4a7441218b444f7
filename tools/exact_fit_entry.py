"""The exact_cmp_fit_n32k_tf32 entry of bench.py alone, three times: python tools/exact_fit_entry.py"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for _ in range(3):
    r = bench.run_exact_fit(torch.device("cuda", 0), 32768, torch.float32)
    print(json.dumps({k: r[k] for k in ("fit_step_seconds", "predict_seconds", "loss")}), flush=True)
