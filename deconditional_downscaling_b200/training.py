"""Training-step drivers: one process per GPU, data-parallel over bag minibatches.

`ElboTrainer.step` is the hot loop body of the reference's variational trainers
(experiments/downscaling/models/variational_cmp.py:163-186,
experiments/swiss_roll/models/vbagg.py:88-101): q = model(x) -> ELBO -> backward -> Adam,
plus - when torch.distributed is initialised - ONE NCCL all-reduce of the flat gradient
buffer (every rank draws its own minibatch; gradients are averaged).
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import functional as F
from . import variational
from .likelihoods import CMPLikelihood, VBaggGaussianLikelihood
from .mlls import BagVariationalELBO


class FlatGrads:
    """All parameter gradients as views into one contiguous buffer, so the data-parallel
    exchange is a single collective (payload ~ M^2 + M d + M + O(10) values)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dtype = self.params[0].dtype
        self.flat = torch.zeros(total, dtype=dtype, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


class ElboTrainer:
    """VariationalCMP (+CMPLikelihood) or VariationalGP (+VBaggGaussianLikelihood) ELBO steps."""

    def __init__(self, model, likelihood, num_data, beta=1.0, lr=0.01, cuda_graph=False):
        self.model, self.likelihood = model, likelihood
        self.elbo = BagVariationalELBO(likelihood, model, num_data=num_data, beta=beta)
        self.params = list(model.parameters()) + list(likelihood.parameters())
        # q(u) must exist before the optimizer / flat buffer are laid out
        model.variational_strategy.ensure_initialized()
        self.grads = FlatGrads(self.params)
        self.opt = torch.optim.Adam(self.params, lr=lr)
        model.train()
        likelihood.train()
        # CUDA-graph mode: forward + backward + gradient all-reduce of a minibatch SHAPE are captured once (third call
        # with that shape) and replayed; the optimizer runs outside the graph, after the factorisation checks
        self.cuda_graph = cuda_graph
        self._graphs, self._graph_calls, self._capture_stream = {}, {}, None
        self.launches_replayed = 0   # libdcgp kernel launches executed through graph replays (host counters do not see them)

    def loss(self, x, z, **kw):
        q = self.model(x)
        if isinstance(self.likelihood, CMPLikelihood):
            kw = self.model.get_elbo_computation_parameters(bags_values=kw["bags_values"],
                                                            extended_bags_values=kw["extended_bags_values"])
        return -self.elbo(variational_dist_f=q, target=z, **kw)

    def _backward_pass(self, x, z, kw):
        """zero grads -> loss -> backward -> all-reduce, optimistic about the factorisations; returns (loss, pending)."""
        self.grads.zero()
        with F.deferred_pd_checks() as pending:
            loss = self.loss(x, z, **kw)
            loss.backward()
        self.grads.allreduce_mean()
        return loss.detach(), pending

    def _checked_pass(self, x, z, kw):
        """psd_safe_cholesky semantics (jitter retries, then NotPSDError) with immediate checks."""
        self.grads.zero()
        loss = self.loss(x, z, **kw)
        loss.backward()
        self.grads.allreduce_mean()
        return loss.detach()

    def step(self, x, z, **kw):
        """One ELBO step on device-resident tensors; returns the (device) loss."""
        if self.cuda_graph:
            return self._graph_step(x, z, kw)
        # optimistic pass: no host read-back between the factorisations, ONE check before the update
        return self._eager_step(x, z, kw)

    # ---- CUDA-graph mode ---------------------------------------------------------------------------------
    def _graph_step(self, x, z, kw):
        """Everything of graph mode - eager warm-up steps, capture, replays, optimizer - runs on ONE dedicated stream:
        autograd's gradient-accumulation nodes are bound to the stream they were first used on, and a node bound to the
        default stream cannot take part in a capture."""
        dev = x.device
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=dev)
        s, cur = self._capture_stream, torch.cuda.current_stream(dev)
        s.wait_stream(cur)
        # the captured schedule also overlaps the K_ZZ factor with the Lambda factorisation (variational.prefetch)
        with torch.cuda.stream(s), variational.side_stream_scope():
            out = self._graph_step_on_stream(x, z, kw)
        cur.wait_stream(s)
        return out

    def _eager_step(self, x, z, kw):
        try:
            loss, pending = self._backward_pass(x, z, kw)
            pending.check()
        except F.NotPSDError:
            loss = self._checked_pass(x, z, kw)
        self.opt.step()
        return loss

    def _graph_step_on_stream(self, x, z, kw):
        tensors = {"x": x, "z": z, **{k: v for k, v in kw.items() if torch.is_tensor(v)}}
        others = {k: v for k, v in kw.items() if not torch.is_tensor(v)}
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(tensors.items())) + \
            tuple((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in sorted(others.items()))
        entry = self._graphs.get(key)
        if entry is None:
            n = self._graph_calls[key] = self._graph_calls.get(key, 0) + 1
            if n <= 2:   # eager warm-up: kernel attributes, look-ahead lanes, allocator
                return self._eager_step(x, z, kw)
            try:
                entry = self._capture(tensors, others)
            except RuntimeError as err:   # e.g. autograd state bound to another stream by earlier eager use
                import warnings
                warnings.warn(f"CUDA-graph capture of the ELBO step failed ({str(err).splitlines()[0]}); staying eager")
                torch.cuda.synchronize(x.device)
                entry = False
            self._graphs[key] = entry
        if entry is False:
            return self._eager_step(x, z, kw)
        static, graph, loss, infos, launches = entry
        for k, v in tensors.items():
            if v is not static[k]:
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        self.launches_replayed += launches
        out = loss.clone()
        if infos is not None and bool(infos.any().item()):
            # a factorisation failed inside the replay: repeat the pass eagerly (retries with jitter or raises)
            out = self._checked_pass(tensors["x"], tensors["z"], {**{k: v for k, v in tensors.items() if k not in ("x", "z")}, **others})
        self.opt.step()
        return out

    def _capture(self, tensors, others):
        """Called on the capture stream."""
        from . import _lib
        dev = tensors["x"].device
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        static = {k: v.clone() for k, v in tensors.items()}
        kw = {**{k: v for k, v in static.items() if k not in ("x", "z")}, **others}
        torch.cuda.synchronize(dev)
        lib = _lib.load()
        graph = torch.cuda.CUDAGraph()
        n0 = lib.dcgp_launch_count()
        with torch.cuda.graph(graph, stream=self._capture_stream):
            loss, pending = self._backward_pass(static["x"], static["z"], kw)
            infos = torch.cat([i.reshape(1) for i in pending.infos]) if pending.infos else None
        launches = int(lib.dcgp_launch_count() - n0)
        return static, graph, loss, infos, launches

    def close(self):
        """Drop the captured graphs (and the NCCL work recorded in them).  Call before
        `torch.distributed.destroy_process_group()`: tearing a communicator down while graphs that hold
        its collectives are alive can block at process exit."""
        if self._capture_stream is not None:
            self._capture_stream.synchronize()
        self._graphs.clear()
        self._graph_calls.clear()
        self.cuda_graph = False

    def step_from_host(self, host_batch: Dict[str, torch.Tensor], device):
        """End-to-end step: pinned host tensors -> device, step, loss back to the host."""
        # persistent device staging buffers (one per input name and shape): no allocator traffic per step
        stage = self.__dict__.setdefault("_stage", {})
        dev = {}
        for k, v in host_batch.items():
            if not torch.is_tensor(v):
                dev[k] = v
                continue
            key = (k, tuple(v.shape), v.dtype)
            buf = stage.get(key)
            if buf is None:
                buf = stage[key] = torch.empty(v.shape, dtype=v.dtype, device=device)
            buf.copy_(v, non_blocking=True)
            dev[k] = buf
        x, z = dev.pop("x"), dev.pop("z")
        return float(self.step(x, z, **dev).item())
