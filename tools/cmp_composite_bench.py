"""dcgp_cmp_elbo_fwd + dcgp_cmp_elbo_bwd at the bench minibatch shape (BASELINE config 5): python tools/cmp_composite_bench.py"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops, _lib
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
N, n, M, d, r = 8192, 1024, 1024, 5, 3
for dt in (torch.float32, torch.float64):
    t = lambda v: v.to(device=dev, dtype=dt)
    k = ops.FlatSpec(["rbf"], [list(range(d))], [t(torch.full((d,), 0.6931))], [t(torch.full((1,), 0.6931))])
    l = ops.FlatSpec(["matern32", "rbf"], [[0, 1], [2]], [t(torch.full((2,), 0.6931)), t(torch.full((1,), 0.6931))],
                     [t(torch.full((1,), 0.6931)), t(torch.full((1,), 0.6931))])
    Z, x, ext, y, z = (t(torch.randn(M, d, generator=g)), t(torch.randn(N, d, generator=g)), t(torch.randn(N, r, generator=g)),
                       t(torch.randn(n, r, generator=g)), t(torch.randn(n, generator=g)))
    m, Ls = t(torch.zeros(M)), t(torch.eye(M))
    noise, ind = t(torch.full((1,), 0.6932)), t(torch.full((1,), 0.6931))
    f = lambda: ops.cmp_elbo(k, l, Z, x, ext, y, z, m, Ls, noise, ind, 1e-4, 10240.0)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    l0 = _lib.load().dcgp_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        elbo, grads, info = f()
    e.record()
    torch.cuda.synchronize()
    ms_eager = s.elapsed_time(e) / 5
    # the same two calls captured once and replayed (one stream, no host synchronisation inside)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        f()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = f()
    graph.replay()
    torch.cuda.synchronize()
    s.record()
    for _ in range(5):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    print(json.dumps({"dtype": str(dt), "ms_fwd_bwd": ms_eager, "ms_fwd_bwd_graph_replay": s.elapsed_time(e) / 5, "elbo_graph": out[0].item(), "launches": (_lib.load().dcgp_launch_count() - l0) / 5,
                      "elbo": elbo.item(), "info": info.tolist(), "workspace_GB": _lib.load().dcgp_cmp_elbo_workspace(N, n, M, d, 1 if dt == torch.float32 else 0) / 1e9}))
