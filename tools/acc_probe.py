"""Accuracy of the bench-scale VariationalCMP ELBO (first step, same minibatch) across precision
modes vs the float64 GPU path and the CPU port.  python tools/acc_probe.py [--cpu]"""
import os, sys, json, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from deconditional_downscaling_b200 import ops

wl = bench.WL
dev = torch.device("cuda", 0)
shard = bench.make_shard(wl, 0)
host = bench.minibatches(wl, shard, 1, seed=99)[0]
out = {}
INF = 1 << 62
for mode, dtype, tc in (("f64", torch.float64, True), ("x3", torch.float32, True), ("fast_tc", torch.float32, True),
                        ("potrf_x3", torch.float32, True), ("solve_x3", torch.float32, True), ("potrf_solve_x3", torch.float32, True),
                        ("gemm_x3", torch.float32, True)):
    ops.TF32_MODE = "x3" if mode == "x3" else ("fast" if mode == "fast_tc" else "auto")
    ops.X3_POTRF_LIMIT = INF if mode in ("potrf_x3", "potrf_solve_x3") else 1024
    ops.X3_SOLVE_LIMIT = INF if mode in ("solve_x3", "potrf_solve_x3") else 1024
    ops.X3_GEMM_LIMIT = INF if mode == "gemm_x3" else 512 ** 3
    ops.USE_TCGEN05 = tc
    tr = bench.build_trainer(wl, dev, dtype=dtype)
    b = {k: v.to(dev).to(dtype) for k, v in host.items()}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = tr.loss(b["x"], b["z"], bags_values=b["bags_values"], extended_bags_values=b["extended_bags_values"])
    loss.backward()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    g = tr.grads.flat.double()
    out[mode] = {"loss": loss.item(), "gnorm": g.norm().item(), "sec": dt}
    if mode == "f64":
        gref = g.clone()
    else:
        out[mode]["grad_relerr"] = ((g - gref).norm() / gref.norm()).item()
        out[mode]["loss_relerr"] = abs(loss.item() - out["f64"]["loss"]) / abs(out["f64"]["loss"])
    print(mode, json.dumps(out[mode]), flush=True)
if "--cpu" in sys.argv:
    from oracle import port
    torch.set_num_threads(os.cpu_count())
    p = {k: v.clone().requires_grad_(True) for k, v in port.make_vcmp_params(wl["M"], wl["d"], wl["r"]).items()}
    loss = port.vcmp_loss(p, host["x"], host["extended_bags_values"], host["bags_values"], host["z"], wl["lbda"], wl["n_bags_total"], wl["beta"])
    print("cpu_port_f32", loss.item(), "relerr vs gpu f64", abs(loss.item() - out["f64"]["loss"]) / abs(out["f64"]["loss"]), flush=True)
