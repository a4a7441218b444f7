"""CPU (gloo, world_size 2) test of the data-parallel plumbing: FlatGrads lays every
gradient into one buffer, one all-reduce averages it, and both ranks take the identical Adam
step — checked against the single-process oracle port on the concatenated minibatches."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deconditional_downscaling_b200.training import FlatGrads
    from oracle import port as oport
    torch.manual_seed(0)
    p = oport.make_vcmp_params(6, 3, 3, dtype=torch.float64, seed=1)
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    flat = FlatGrads(list(params.values()))
    opt = torch.optim.Adam(list(params.values()), lr=0.05)
    g = torch.Generator().manual_seed(10 + rank)       # every rank draws its own minibatch
    x, ext = torch.randn(20, 3, generator=g, dtype=torch.float64), torch.randn(20, 3, generator=g, dtype=torch.float64)
    y, z = torch.randn(4, 3, generator=g, dtype=torch.float64), torch.randn(4, generator=g, dtype=torch.float64)
    flat.zero()
    loss = oport.vcmp_loss(params, x, ext, y, z, 1e-2, 4)
    loss.backward()
    local = flat.flat.clone()
    flat.allreduce_mean()
    opt.step()
    out[rank] = (local, flat.flat.clone(), {k: v.detach().clone() for k, v in params.items()})
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    l0, a0, p0 = out[0]
    l1, a1, p1 = out[1]
    assert not torch.allclose(l0, l1)                          # ranks really had different gradients
    torch.testing.assert_close(a0, (l0 + l1) / 2, rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(a0, a1, rtol=0, atol=0)
    for k in p0:                                               # identical parameters after the step
        torch.testing.assert_close(p0[k], p1[k], rtol=0, atol=0)


def test_flat_grads_views_alias_parameter_grads():
    from deconditional_downscaling_b200.training import FlatGrads
    a = torch.nn.Parameter(torch.zeros(3, 2))
    b = torch.nn.Parameter(torch.zeros(4))
    f = FlatGrads([a, b])
    (a.sum() * 2 + (b * torch.arange(4.0)).sum()).backward()
    assert f.slice_of(a).tolist() == [2.0] * 6 and f.slice_of(b).tolist() == [0.0, 1.0, 2.0, 3.0]
    assert f.flat.tolist() == [0.0, 1.0, 2.0, 3.0] + [2.0] * 6          # the largest gradient sits at the end of the buffer
    f.zero()
    assert a.grad.abs().sum() == 0 and b.grad.abs().sum() == 0
    assert f.nbytes == 40


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deconditional_downscaling_b200.parallel import sharded_rows_matmul
    torch.manual_seed(0)
    K = torch.randn(11, 7, dtype=torch.float64)      # 11 rows over 2 ranks: ragged blocks (6, 5)
    W = torch.randn(7, 3, dtype=torch.float64)
    res = sharded_rows_matmul(lambda lo, hi: K[lo:hi] @ W, 11, 3, torch.float64, "cpu")
    out[rank] = (res, K @ W)
    dist.destroy_process_group()


def test_row_block_sharded_product_two_ranks():
    """Row-block sharding + one all-gather reproduces the unsharded product on every rank."""
    from deconditional_downscaling_b200.parallel import row_block
    assert [row_block(11, r, 2) for r in range(2)] == [(0, 6), (6, 11)]
    assert [row_block(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        res, ref = out[r]
        torch.testing.assert_close(res, ref, rtol=1e-13, atol=1e-13)


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """bench.py --impl reference launched as the driver launches it for N > 1: rank 0 alone runs the CPU arm and prints
    the JSON line, the other ranks leave with exit code 0 and no output."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(root, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--small", "--steps", "1", "--warmup", "0"],
                       cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["metric"] == "elbo_steps_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


# ---------------------------------------------------------------------------------------------------------------
# step protocol of training.DataParallelStepper (collective failure decision BEFORE any update, rank-0 broadcast of the
# initial state) on two gloo ranks, with a toy compute in place of the CUDA kernels
def _protocol_worker(rank, world, port, out, fail_rank, fail_step, retry_fails, overlap=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deconditional_downscaling_b200 import functional as F
    from deconditional_downscaling_b200.training import DataParallelStepper, FlatGrads, sync_module_state

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(100 + rank)                         # every rank starts DIFFERENT ...
            self.w = torch.nn.Parameter(torch.randn(5, dtype=torch.float64))
            self.big = torch.nn.Parameter(torch.randn(4, 5, dtype=torch.float64))   # exchanged early, by the hook
            self.register_buffer("freq", torch.randn(3, dtype=torch.float64))   # ... incl. a lazily drawn buffer

    class Stepper(DataParallelStepper):
        def __init__(self):
            self.m = Toy()
            sync_module_state([self.m])                           # ... and is made identical here
            self.grads = FlatGrads([self.m.w, self.m.big])
            if overlap:
                self.grads.overlap_largest()
            self.pool = F.InfoPool("cpu")
            self.opt = torch.optim.Adam([self.m.w, self.m.big], lr=0.1)
            self.t, self.checked, self.updates = 0, 0, 0

        def _loss(self, batch):
            return ((self.m.w * batch).sum() - 1.0) ** 2 + ((self.m.big @ batch).sum() - 0.5) ** 2 + self.m.freq.sum() * 0

        def optimistic_pass(self, batch):
            bad = rank == fail_rank and self.t == fail_step
            self.pool.next().fill_(3 if bad else 0)
            loss = self._loss(batch) * (float("nan") if bad else 1.0)    # a failed factorisation poisons the gradient
            loss.backward()
            return loss.detach()

        def checked_pass(self, batch):
            self.checked += 1
            if retry_fails and rank == fail_rank:
                raise F.NotPSDError("retry failed")
            loss = self._loss(batch)
            loss.backward()
            return loss.detach()

        def update(self, veto):
            if veto is not None and bool(veto.any()):
                return
            self.grads.flat.div_(world)
            self.opt.step()
            self.updates += 1
            self.grads.zero()

    s = Stepper()
    start = torch.cat([s.m.w.detach().reshape(-1), s.m.big.detach().reshape(-1)]), s.m.freq.clone()
    g = torch.Generator().manual_seed(7 + rank)
    raised = False
    for t in range(4):
        s.t = t
        batch = torch.randn(5, generator=g, dtype=torch.float64)
        try:
            s.resolve(batch, s.device_step(batch))
        except F.NotPSDError:
            raised = True
            break
    final = torch.cat([s.m.w.detach().reshape(-1), s.m.big.detach().reshape(-1)])
    out[rank] = (start, final, s.checked, s.updates, raised, bool(torch.isfinite(final).all()))
    dist.destroy_process_group()


def _run_protocol(fail_rank, fail_step, retry_fails, overlap=True):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_protocol_worker, args=(2, _free_port(), out, fail_rank, fail_step, retry_fails, overlap), nprocs=2, join=True)
    return out[0], out[1]


def test_replicas_start_identical_and_stay_identical_through_a_one_rank_failure():
    """(1) parameters AND buffers are broadcast from rank 0 at construction; (2) a factorisation failing on
    ONE rank makes BOTH ranks repeat the step (no rank applies the poisoned gradient, no unmatched collective, no hang)
    and the replicas hold identical finite parameters afterwards."""
    r0, r1 = _run_protocol(fail_rank=1, fail_step=2, retry_fails=False)
    (w0, f0), (w1, f1) = r0[0], r1[0]
    assert torch.equal(w0, w1) and torch.equal(f0, f1)
    assert torch.equal(r0[1], r1[1]) and r0[5] and r1[5]
    assert (r0[2], r1[2]) == (1, 1)              # both ranks took the checked pass, exactly once
    assert (r0[3], r1[3]) == (4, 4)              # 3 optimistic updates + 1 checked update each
    assert not r0[4] and not r1[4]


def test_without_failures_no_rank_takes_the_checked_pass():
    r0, r1 = _run_protocol(fail_rank=-1, fail_step=-1, retry_fails=False)
    assert (r0[2], r1[2]) == (0, 0) and (r0[3], r1[3]) == (4, 4) and torch.equal(r0[1], r1[1])


def test_early_allreduce_of_the_largest_gradient_changes_nothing():
    """FlatGrads.overlap_largest(): the big slice is exchanged by a post-accumulate-grad hook during the backward pass, the
    end-of-step collective carries only the head - same parameters as one collective over the whole buffer, with and
    without a failing rank (the checked pass runs with the hook disabled)."""
    for fail in ((-1, -1), (1, 2)):
        a0, a1 = _run_protocol(fail[0], fail[1], False, overlap=True)
        b0, b1 = _run_protocol(fail[0], fail[1], False, overlap=False)
        assert torch.equal(a0[1], a1[1]) and torch.equal(b0[1], b1[1])
        torch.testing.assert_close(a0[1], b0[1], rtol=1e-14, atol=1e-15)


def test_unrecoverable_failure_on_one_rank_raises_on_all_ranks():
    r0, r1 = _run_protocol(fail_rank=1, fail_step=1, retry_fails=True)
    assert r0[4] and r1[4]                        # NotPSDError everywhere, nobody left waiting in a collective
    assert torch.equal(r0[1], r1[1])              # and nobody applied a half-exchanged update
    assert (r0[3], r1[3]) == (1, 1)
