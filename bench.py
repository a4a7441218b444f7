#!/usr/bin/env python
"""Headline benchmark of the B200-native deconditional-GP hot path.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun launches N>1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm, rank 0 only)

Metric (BASELINE.json): "ELBO steps/s" on config 5 — synthetic VariationalCMP, 1,048,576
individuals / 10,240 bags over 8 shards, M = 1024 inducing points; each rank's step is one
bag minibatch (1024 bags + 8192 CME individuals, individuals-noise ELBO) = forward +
backward + gradient all-reduce + Adam.  Weak scaling: per-GPU work is fixed, `value` is the
number of minibatch steps the whole job completes per second.  At N = 1 the line also carries
`exact_cmp_n32k`: the FP64 Gram + Cholesky + solve composite of config 4 at N = 32768.

Prints ONE JSON line on rank 0 (contract in the task statement).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

# ------------------------------------------------------------------ workload definition
WL = dict(name="variational_cmp_1M_points_10k_bags_M1024", n_total=1_048_576, n_bags_total=10_240, shards=8,
          bags_per_step=1024, individuals_per_step=8192, M=1024, d=5, r=3, lbda=1e-4, lr=0.01, beta=1.0)
SMALL = dict(WL, name="variational_cmp_small", n_total=65_536, n_bags_total=640, bags_per_step=128,
             individuals_per_step=1024, M=128)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "MEASURED_PEAKS.json"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def make_shard(wl, rank, dtype=torch.float32):
    """Synthetic shard of rank `rank`: individuals x ~ N(0, I_d) grouped in contiguous bags,
    bag covariates y ~ N(0, I_r), per-individual bag covariates y~ = y_bag + 0.3 eps, targets z."""
    g = torch.Generator().manual_seed(1234 + rank)
    n_bags = wl["n_bags_total"] // wl["shards"]
    per_bag = wl["n_total"] // wl["n_bags_total"]          # 102 individuals per bag (1,044,480 used of 1,048,576)
    n_ind = n_bags * per_bag
    x = torch.randn(n_ind, wl["d"], generator=g, dtype=dtype)
    y = torch.randn(n_bags, wl["r"], generator=g, dtype=dtype)
    ext = y.repeat_interleave(per_bag, dim=0) + 0.3 * torch.randn(n_ind, wl["r"], generator=g, dtype=dtype)
    w = torch.randn(wl["d"], generator=g, dtype=dtype)
    f = torch.sin(x @ w)
    z = f.view(n_bags, per_bag).mean(1) + 0.05 * torch.randn(n_bags, generator=g, dtype=dtype)
    return x, ext, y, z


def minibatches(wl, shard, n_batches, seed):
    """Reference batch iterator (experiments/downscaling/models/variational_cmp.py:120-136):
    a random subset of bags (y, z) + an independent multinomial sample of CME individuals."""
    x, ext, y, z = shard
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        bi = torch.randperm(y.size(0), generator=g)[:wl["bags_per_step"]]
        ii = torch.randperm(x.size(0), generator=g)[:wl["individuals_per_step"]]
        out.append({"x": x[ii].contiguous(), "extended_bags_values": ext[ii].contiguous(),
                    "bags_values": y[bi].contiguous(), "z": z[bi].contiguous()})
    return out


def batch_bytes(b):
    return sum(v.numel() * v.element_size() for v in b.values())


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ GEMM instrumentation (roofline)
class GemmProbe:
    """CUDA-event pairs on the launching stream around every call that runs the tensor-core
    contraction kernel gemm_kernel<T>: dcgp_gemm itself and the blocked drivers built on it
    (dcgp_potrf, dcgp_trsm, dcgp_potrs).  Algorithmic flops from the shapes: 2 m n k (m n k for
    lower-only), n^3/3 (potrf), n^2 m (trsm), 2 n^2 m (potrs, as two triangular solves)."""

    def __init__(self, ops):
        self.ops, self.rec, self.on = ops, [], False
        self.orig = {name: getattr(ops, name) for name in ("gemm", "potrf_", "trsm_", "potrs_")}

    def _wrap(self, name, flops_fn):
        probe, orig = self, self.orig[name]

        def wrapped(*a, **k):
            if not probe.on:
                return orig(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = orig(*a, **k)
            e.record()
            probe.rec.append((s, e, flops_fn(*a, **k), name, a[0].dtype))
            return out
        setattr(self.ops, name, wrapped)

    def install(self):
        def gemm_flops(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None, lower=False, no_splitk=False):
            m = A.size(1) if transa else A.size(0)
            k = A.size(0) if transa else A.size(1)
            n = B.size(0) if transb else B.size(1)
            return (1.0 if lower else 2.0) * m * n * k
        self._wrap("gemm", gemm_flops)
        self._wrap("potrf_", lambda A, info=None: A.size(0) ** 3 / 3.0)
        self._wrap("trsm_", lambda L, B, trans=False, plan=None: float(B.size(0)) ** 2 * B.size(1))
        self._wrap("potrs_", lambda L, B, plan=None: 2.0 * float(B.size(0)) ** 2 * B.size(1))

    def summary(self):
        torch.cuda.synchronize()
        t = sum(r[0].elapsed_time(r[1]) for r in self.rec) * 1e-3
        fl = sum(r[2] for r in self.rec)
        by = {}
        for s, e, f, name, _ in self.rec:
            d = by.setdefault(name, [0, 0.0, 0.0])
            d[0] += 1
            d[1] += s.elapsed_time(e) * 1e-3
            d[2] += f
        return len(self.rec), t, fl, {k: {"calls": v[0], "ms": 1e3 * v[1], "tflops": v[2] / v[1] / 1e12 if v[1] > 0 else None} for k, v in by.items()}

    def tcgen05_products(self, min_flops=2e9):
        """Direct FP32 dcgp_gemm calls large enough to be one gemm_tc_kernel launch each (gemm.cu dispatch:
        m, n >= 128, k >= 64, >= 32 tiles): launches, seconds, algorithmic flops."""
        sel = [r for r in self.rec if r[3] == "gemm" and r[4] == torch.float32 and r[2] >= min_flops]
        return len(sel), sum(r[0].elapsed_time(r[1]) for r in sel) * 1e-3, sum(r[2] for r in sel)


# ------------------------------------------------------------------ our arm
def build_trainer(wl, device, dtype=torch.float32, cuda_graph=False):
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M
    from deconditional_downscaling_b200.training import ElboTrainer
    inv_sp1 = math.log(math.e - 1.0)
    k = K.ScaleKernel(K.RBFKernel(ard_num_dims=wl["d"]))
    k.base_kernel.initialize(raw_lengthscale=torch.full((wl["d"],), inv_sp1))
    l0 = K.MaternKernel(nu=1.5, ard_num_dims=2, active_dims=[0, 1])
    l0.initialize(raw_lengthscale=torch.full((2,), inv_sp1))
    l1 = K.RBFKernel(ard_num_dims=wl["r"] - 2, active_dims=list(range(2, wl["r"])))
    l1.initialize(raw_lengthscale=torch.full((wl["r"] - 2,), inv_sp1))
    l = K.ScaleKernel(l0) + K.ScaleKernel(l1)
    g = torch.Generator().manual_seed(0)
    Z = torch.randn(wl["M"], wl["d"], generator=g)
    model = M.VariationalCMP(inducing_points=Z, individuals_mean=Mn.ZeroMean(), individuals_kernel=k, bag_kernel=l,
                             lbda=wl["lbda"], use_individuals_noise=True).to(device).to(dtype)
    lik = L.CMPLikelihood(use_individuals_noise=True).to(device).to(dtype)
    vd = model.variational_strategy._variational_distribution
    with torch.no_grad():   # deterministic q(u) init (m = 0, L_S = I) on every rank
        vd.variational_mean.zero_()
        vd.chol_variational_covar.copy_(torch.eye(wl["M"]))
    model.variational_strategy.variational_params_initialized.fill_(1)
    return ElboTrainer(model, lik, num_data=wl["n_bags_total"], beta=wl["beta"], lr=wl["lr"], cuda_graph=cuda_graph)


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdcgp has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    wl = SMALL if args.small else WL
    from deconditional_downscaling_b200 import ops, _lib
    lib = _lib.load()
    shard = make_shard(wl, rank % wl["shards"])
    K_, W_ = args.steps, max(args.warmup, 3)
    host = minibatches(wl, shard, K_ + W_, seed=99 + rank)
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    devb = [{k: v.to(device) for k, v in b.items()} for b in host]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn):
        """W_ warm-up steps, then K_ steps between events on the current stream; max over ranks."""
        first = None
        for i in range(W_):
            l_ = step_fn(i)
            if i == 0:
                first = float(l_.item()) if torch.is_tensor(l_) else float(l_)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(W_, W_ + K_):
            last = step_fn(i)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), first, (float(last.item()) if torch.is_tensor(last) else float(last))

    # ---- eager pass with CUDA events around every tensor-core contraction call: the roofline numbers.  (Events cannot
    # be recorded around individual kernels of a graph replay, so the same steps are issued eagerly once.)
    probe = GemmProbe(ops)
    probe.install()
    eager = build_trainer(wl, device)

    def eager_step(i):
        b = dict(devb[i])
        probe.on = i >= W_
        return eager.step(b.pop("x"), b.pop("z"), **b)
    t_eager, _, _ = timed(eager_step)
    probe.on = False
    del eager
    torch.cuda.empty_cache()

    # ---- device-resident arm: the step as a user of the package runs it (CUDA-graph replay unless --no-graph) ---------
    trainer = build_trainer(wl, device, cuda_graph=not args.no_graph)

    def dev_step(i):
        b = dict(devb[i])
        return trainer.step(b.pop("x"), b.pop("z"), **b)
    for i in range(W_):          # warm-up here (capture happens on the third call), timed region below
        l_ = dev_step(i)
        if i == 0:
            first_loss = float(l_.item())   # ELBO of minibatch 0 at the initial parameters (parity check below)
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    sync_all()
    l0 = lib.dcgp_launch_count() + trainer.launches_replayed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W_, W_ + K_):
        loss = dev_step(i)
    e1.record()
    sync_all()
    launches = lib.dcgp_launch_count() + trainer.launches_replayed - l0
    t_dev = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev = t_dev.item()
    clk = clocks.stop() if rank == 0 else None
    final_loss = float(loss.item())
    n_gemm, t_gemm, fl_gemm, gemm_by = probe.summary()
    n_tc, t_tc, fl_tc = probe.tcgen05_products()
    # ---- end-to-end arm: pinned host -> device copies and loss read-back inside the timed region
    for i in range(W_):
        trainer.step_from_host(host[i], device)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W_, W_ + K_):
        trainer.step_from_host(host[i], device)
    e1.record()
    sync_all()
    t_e2e = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = t_e2e.item()

    pk = peaks()
    tf32_peak = pk["bf16_tflops_sustained"] / 2 if pk.get("bf16_tflops_sustained") else pk["bf16_tflops"] / 2
    out = {
        "metric": "elbo_steps_per_s", "value": world * K_ / t_dev, "unit": "minibatch ELBO steps/s (whole job)",
        "n_gpus": world, "steps": K_, "warmup": W_, "ms_per_step": 1e3 * t_dev / K_, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 storage; products on single-pass TF32 (tcgen05), factorisations/solves on 3xTF32; f64 island for K_ZZ",
        "data": "synthetic",
        "config": {"workload": wl["name"], "bags_per_step_per_gpu": wl["bags_per_step"],
                   "cme_individuals_per_step_per_gpu": wl["individuals_per_step"], "inducing_points": wl["M"],
                   "d": wl["d"], "r": wl["r"], "lbda": wl["lbda"], "likelihood": "CMPLikelihood(use_individuals_noise=True)",
                   "step": "forward + backward + flat-gradient all-reduce + Adam", "parallelism": f"dp{world}",
                   "step_mode": "eager" if args.no_graph else "CUDA-graph replay of forward + backward + all-reduce (captured once, "
                                "third warm-up step), factorisation check, eager Adam",
                   "l2": "per-step working set (Lambda 8192^2 f32 = 268 MB plus K_xx) exceeds the 126 MB L2; "
                         "a fresh minibatch every step"},
        "e2e": {"value": world * K_ / t_e2e, "unit": "minibatch ELBO steps/s (whole job)",
                "h2d_bytes_per_step": batch_bytes(host[0]), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "final_loss": final_loss,
        "value_eager": world * K_ / t_eager,
        # dominant throughput kernel (profiles/r01_launches_bench_step_v15_summary.csv: gemm_tc_kernel 51% of the
        # kernel time of a step; the single-CTA potrf_step_kernel chain, 25%, is latency- not roofline-bound)
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05.mma kind::tf32, TMA, TMEM): the direct dcgp_gemm "
                     "products of the timed steps that run as one launch of it each, CUDA events around every launch",
                     "achieved": fl_tc / t_tc / 1e12 if t_tc > 0 else None, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": (fl_tc / t_tc / 1e12) / tf32_peak if t_tc > 0 else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step's most frequent large product
                     # (K_xx[8192 x 8192] @ A[8192 x 1024], 335 MB compulsory) from the ncu --set full capture
                     # profiles/r01_ncu_tcgemm_v2_8192x1024x8192_raw.csv
                     "traffic": 425892096, "traffic_shape": "8192x1024x8192 NN",
                     "peak_source": f"{pk['source']}: sustained bf16 cuBLAS rate / 2 (TF32 issues at half the bf16 MMA rate)",
                     "launches": n_tc, "avg_launch_ms": 1e3 * t_tc / n_tc if n_tc else None,
                     "share_of_step": t_tc / t_eager if t_eager > 0 else None,
                     "path": {"what": "all tensor-core contractions of the step (dcgp_gemm, and the GEMMs inside dcgp_potrf / "
                              "dcgp_trsm_planned incl. their serial leaf chains): algorithmic flops / event time",
                              "achieved": fl_gemm / t_gemm / 1e12 if t_gemm > 0 else None,
                              "frac": (fl_gemm / t_gemm / 1e12) / tf32_peak if t_gemm > 0 else None,
                              "calls": n_gemm, "share_of_step": t_gemm / t_eager if t_eager > 0 else None, "by_call": gemm_by,
                              "measured_on": "the eager issue of the same steps (value_eager)"}},
        "allreduce_bytes_per_step": trainer.grads.nbytes,
    }
    if rank == 0:
        out["clocks"] = clk
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(wl, host, steps=2)
            c0 = out["cpu_baseline"].pop("first_loss")
            out["parity"] = {"what": "-ELBO of minibatch 0 at the initial parameters: this run vs the CPU oracle port",
                             "gpu": first_loss, "cpu_oracle": c0,
                             "rel_diff": abs(first_loss - c0) / abs(c0) if first_loss is not None else None,
                             "tolerance": 1e-3}
        if world == 1 and not args.no_exact:
            out["exact_cmp_n32k"] = run_exact(device, 4096 if args.small else 32768)
            out["vbagg_1024bags"] = run_vbagg(device)
        print(json.dumps(out), flush=True)
    if world > 1:
        # Orderly end of a multi-rank run: everyone has finished its device work, the graphs that hold captured NCCL
        # all-reduces are dropped first, and the process leaves without the communicator teardown (which was seen to
        # block for minutes after graph capture, long after rank 0 had printed its line).
        sync_all()
        trainer.close()
        del trainer
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------ ExactCMP N = 32k composite (N = 1 only)
def run_exact(device, N):
    from deconditional_downscaling_b200 import exact
    dt = torch.float64
    g = torch.Generator().manual_seed(7)
    n = N // 100
    x = torch.randn(N, 3, generator=g, dtype=dt)
    x = (x - x.mean(0)) / x.std(0)
    ybag = torch.randn(n, 3, generator=g, dtype=dt)
    bag_of = torch.arange(N) * n // N
    ext = ybag[bag_of] + 0.3 * torch.randn(N, 3, generator=g, dtype=dt)     # r = 3 continuous variant
    z = torch.randn(n, generator=g, dtype=dt)
    x, ext, ybag, z = (t.to(device) for t in (x, ext, ybag, z))
    # measured FP64 tensor peak: cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json has no FP64 entry)
    A = torch.randn(8192, 8192, dtype=dt, device=device)
    B = torch.randn(8192, 8192, dtype=dt, device=device)
    C = torch.empty_like(A)
    best = 1e9
    for i in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(A, B, out=C)
        e.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, s.elapsed_time(e) * 1e-3)
    fp64_peak = 2 * 8192 ** 3 / best / 1e12
    del A, B, C
    exact.exact_cmp_phases(x[:4096], ext[:4096], ybag[:40], z[:40], x[:4096], 0.01)   # warm-up: kernels
    res = exact.exact_cmp_phases(x, ext, ybag, z, x, 0.01)                            # warm-up: allocator at full size
    del res
    torch.cuda.synchronize()
    timer = exact.PhaseTimer(True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    res = exact.exact_cmp_phases(x, ext, ybag, z, x, 0.01, timer=timer)
    e.record()
    torch.cuda.synchronize()
    total = s.elapsed_time(e) * 1e-3
    ph = timer.seconds()
    fl = exact.exact_cmp_flops(N, n, 3, N)
    tot_fl = sum(fl.values())
    return {"metric": "gram_cholesky_solve_tflops", "N": N, "n_bags": n, "dtype": "f64",
            "seconds": total, "tflops": tot_fl / total / 1e12, "algorithmic_flops": tot_fl,
            "phases_ms": {k: 1e3 * v for k, v in ph.items()},
            "potrf_tflops": fl["potrf"] / ph["potrf"] / 1e12, "potrs_tflops": fl["potrs"] / ph["potrs"] / 1e12,
            "gram_K_gbs": N * N * 8 / ph["gram_K"] / 1e9,
            "roofline": {"bound": "tensor", "kernel": "gemm_kernel<double> (DMMA) inside dcgp_potrf",
                         "achieved": fl["potrf"] / ph["potrf"] / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": fl["potrf"] / ph["potrf"] / 1e12 / fp64_peak,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (best of 3)"},
            # the Gram build is bound by the N^2 * 8 bytes it writes; FP64 exp (about 25 DFMA per entry) is the co-bound
            "gram_roofline": {"bound": "hbm", "kernel": "gram_fwd_kernel<double>", "achieved": N * N * 8 / ph["gram_K"] / 1e9,
                              "peak": peaks()["hbm_gbs"], "unit": "GB/s",
                              "frac": N * N * 8 / ph["gram_K"] / 1e9 / peaks()["hbm_gbs"], "peak_source": peaks()["source"]},
            "potrf_info": int(res["info"].item()), "mll": float(res["mll"].item())}


# ------------------------------------------------------------------ VBagg ELBO step at config-5 scale (N = 1 only)
def run_vbagg(device, steps=10, warmup=3):
    """1024 bags x 102 individuals per step (104 448 rows), M = 1024, d = 5, FP32: forward + backward + Adam.
    The bag-summed cross Gram evaluates M x N_b kernel values per pass without storing K_Zx."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M
    from deconditional_downscaling_b200.training import ElboTrainer
    g = torch.Generator().manual_seed(5)
    nb, per, Mz, d = 1024, 102, 1024, 5
    sizes = [per] * nb
    model = M.VariationalGP(inducing_points=torch.randn(Mz, d, generator=g), mean_module=Mn.ZeroMean(),
                            covar_module=K.ScaleKernel(K.RBFKernel(ard_num_dims=d)))
    lik = L.VBaggGaussianLikelihood()
    model, lik = model.to(device), lik.to(device)
    tr = ElboTrainer(model, lik, num_data=10240, beta=1.0, lr=0.01)
    batches = [(torch.randn(nb * per, d, generator=g).to(device), torch.randn(nb, generator=g).to(device))
               for _ in range(steps + warmup)]
    for i in range(warmup):
        tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(warmup, warmup + steps):
        loss = tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    evals = float(Mz) * nb * per
    return {"metric": "vbagg_elbo_steps_per_s", "value": 1e3 / ms, "ms_per_step": ms, "bags_per_step": nb,
            "individuals_per_step": nb * per, "inducing_points": Mz, "dtype": "f32 (f64 K_ZZ factor)",
            "cross_gram_kernel_evals_per_step": evals, "final_loss": float(loss.item()),
            "note": "bag-summed cross Gram forward + backward = 2 x M x N_b kernel evaluations per step, K_Zx never stored"}


# ------------------------------------------------------------------ CPU arms (oracle port; the only place oracle/ runs)
def cpu_steps(wl, host, steps, warmup=0):
    from oracle import port
    torch.set_num_threads(os.cpu_count() or 1)
    tr = port.CpuVcmpTrainer(port.make_vcmp_params(wl["M"], wl["d"], wl["r"]), lr=wl["lr"])
    args = lambda b: (b["x"], b["extended_bags_values"], b["bags_values"], b["z"], wl["lbda"], wl["n_bags_total"], wl["beta"])
    for i in range(warmup):
        tr.step(*args(host[i % len(host)]))
    t0 = time.perf_counter()
    first = None
    for i in range(steps):
        loss = tr.step(*args(host[(warmup + i) % len(host)]))
        if first is None:
            first = float(loss)
    dt = time.perf_counter() - t0
    return dt, float(loss), first


def cpu_baseline(wl, host, steps=2):
    dt, loss, first = cpu_steps(wl, host, steps)
    return {"first_loss": first, "value": steps / dt, "unit": "minibatch ELBO steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{steps} full minibatch steps of the same workload (oracle/port.py, torch CPU, all host threads)",
            "final_loss": loss}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = SMALL if args.small else WL
    shard = make_shard(wl, 0)
    host = minibatches(wl, shard, args.steps + args.warmup, seed=99)
    dt, loss, _ = cpu_steps(wl, host, args.steps, warmup=min(args.warmup, 1))
    v = args.steps / dt
    unit = "minibatch ELBO steps/s (whole job)"
    print(json.dumps({
        "impl": "reference", "metric": "elbo_steps_per_s", "value": v, "unit": unit, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 island for K_ZZ)", "data": "synthetic",
        "config": {"workload": wl["name"], "note": "reference algorithm restated on torch CPU (oracle/port.py): gpytorch 1.4.0 "
                   "is not installable offline, so the unmodified reference cannot run"},
        "cpu_baseline": {"value": v, "unit": unit, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} full minibatch steps on rank 0 only"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "final_loss": loss}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="reduced sizes (smoke / CI)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-exact", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
