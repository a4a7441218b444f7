"""Training-step drivers: one process per GPU, data-parallel over bag minibatches.

`ElboTrainer.step` is the hot loop body of the reference's variational trainers
(experiments/downscaling/models/variational_cmp.py:163-186,
experiments/swiss_roll/models/vbagg.py:88-101): q = model(x) -> ELBO -> backward -> Adam,
plus - when torch.distributed is initialised - ONE NCCL all-reduce of the flat gradient
buffer (every rank draws its own minibatch; gradients are averaged).

Step protocol (`DataParallelStepper`, the same on every rank, no rank-local control flow between collectives):
  1. optimistic pass: forward + backward with every Cholesky failure word written into one device pool
     (functional.InfoPool) - no host read-back;
  2. MAX all-reduce of the pool (a few dozen ints): afterwards every rank holds the union of all ranks' failures;
  3. SUM all-reduce of the flat gradient buffer - always, by every rank;
  4. fused device-side Adam over the flat parameter buffer, predicated on the pool (dcgp_adam_step `veto`): if any
     rank failed, no rank updates;
  5. ONE host read of the pool; if it is non-zero EVERY rank repeats the minibatch with immediate checks and
     psd_safe_cholesky's jitter retries (collectively: a rank whose retry fails reports it through a second flag
     all-reduce and all ranks raise NotPSDError together).
Steps 1-4 are what `cuda_graph=True` captures and replays.
"""
import os
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import functional as F
from . import variational
from .likelihoods import CMPLikelihood, VBaggGaussianLikelihood
from .mlls import BagVariationalELBO


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class FlatGrads:
    """All parameter gradients as views into one contiguous buffer, so the data-parallel
    exchange is a single collective (payload ~ M^2 + M d + M + O(10) values).

    Layout: parameters in the given order, except that the LARGEST one (chol_variational_covar: M^2 of the
    M^2 + M d + M + O(10) values) sits at the end - `overlap_largest()` then all-reduces that slice on a side stream
    the moment autograd has finished accumulating it (a post-accumulate-grad hook), under the rest of the backward
    pass, and the collective at the end of the step only carries the small head of the buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dtype = self.params[0].dtype
        if any(p.dtype != dtype for p in self.params):
            raise TypeError("FlatGrads needs parameters of one dtype (move the model with .to(dtype) first)")
        self.flat = torch.zeros(total, dtype=dtype, device=self.params[0].device)
        self.big = max(self.params, key=lambda p: p.numel())
        self.offset, off = {}, 0
        for p in [q for q in self.params if q is not self.big] + [self.big]:
            self.offset[id(p)] = off
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.early_enabled = False      # set by overlap_largest()
        self._early_done, self._side, self._hook = False, None, None

    def slice_of(self, p, buf=None):
        off = self.offset[id(p)]
        return (self.flat if buf is None else buf)[off:off + p.numel()]

    def zero(self):
        self.flat.zero_()

    def overlap_largest(self):
        """Install the early all-reduce of the largest gradient (multi-rank runs only)."""
        if self._hook is not None or _world() == 1:
            return
        big, piece = self.big, self.slice_of(self.big)

        def hook(_):
            if not self.early_enabled or _world() == 1:
                return
            if piece.is_cuda:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=piece.device)
                self._side.wait_stream(torch.cuda.current_stream(piece.device))
                with torch.cuda.stream(self._side):
                    dist.all_reduce(piece, op=dist.ReduceOp.SUM)
            else:
                dist.all_reduce(piece, op=dist.ReduceOp.SUM)
            self._early_done = True

        self._hook = big.register_post_accumulate_grad_hook(hook)
        self.early_enabled = True

    def allreduce_sum(self):
        if _world() == 1:
            return
        if self._early_done:        # the largest slice is already on its way (or done): only the head remains
            self._early_done = False
            head = self.flat[:self.offset[id(self.big)]]
            if head.numel():
                dist.all_reduce(head, op=dist.ReduceOp.SUM)
            if self._side is not None:
                torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    def allreduce_mean(self):
        if _world() > 1:
            self.allreduce_sum()
            self.flat.div_(_world())

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


class FlatParams(FlatGrads):
    """FlatGrads + the parameters themselves re-laid as views of one buffer (same layout), so the optimiser update is
    one launch over two flat arrays and the initial rank-0 broadcast one collective."""

    def __init__(self, params):
        super().__init__(params)
        self.values = torch.empty_like(self.flat)
        with torch.no_grad():
            for p in self.params:
                view = self.slice_of(p, self.values).view_as(p)
                view.copy_(p)
                p.data = view


def sync_module_state(modules: Iterable[torch.nn.Module], flat_params: Optional[torch.Tensor] = None, src: int = 0):
    """Broadcast parameters and buffers from rank `src`: replicas must start identical (q(u)'s mean is initialised
    with 1e-3 * randn and RFF frequencies are drawn lazily - both per process)."""
    if _world() == 1:
        return
    if flat_params is not None:
        dist.broadcast(flat_params, src)
    else:
        for m in modules:
            for p in m.parameters():
                dist.broadcast(p.data, src)
    for m in modules:
        for b in m.buffers():
            if b.numel():
                dist.broadcast(b, src)


def materialize_lazy_buffers(model):
    """RFF kernels draw their frequency buffers at first evaluation (kernels.RFFKernel._terms); do it now so that
    they exist on every rank before the broadcast."""
    vs = getattr(model, "variational_strategy", None)
    if vs is None:
        return
    kernel, _ = vs._prior_modules()
    kernel._terms(vs.inducing_points.size(-1), None)


class FusedAdam:
    """torch.optim.Adam arithmetic (default betas / eps, no weight decay, no amsgrad) as ONE libdcgp launch over the
    flat buffers, predicated on device failure words.  state_dict() is laid out like torch.optim.Adam's so that the
    reference's checkpoints ('optimizer' entry of state.pt, experiments/*/models/*.py) keep their format."""

    def __init__(self, flat: FlatParams, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.flat = flat
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)
        self.param_groups = [dict(self.defaults, params=flat.params)]
        self.exp_avg = torch.zeros_like(flat.values)
        self.exp_avg_sq = torch.zeros_like(flat.values)
        self.step_count = torch.zeros(1, dtype=torch.float64, device=flat.values.device)

    def step(self, veto: Optional[torch.Tensor] = None, grad_scale: float = 1.0, zero_grads: bool = False):
        from . import ops
        g = self.param_groups[0]
        ops.adam_step(self.flat.values, self.flat.flat, self.exp_avg, self.exp_avg_sq, self.step_count, g["lr"],
                      g["betas"][0], g["betas"][1], g["eps"], grad_scale, veto, zero_grads)

    def zero_grad(self, set_to_none=False):
        self.flat.zero()

    def state_dict(self):
        state = {}
        for i, p in enumerate(self.flat.params):
            state[i] = {"step": self.step_count.clone().reshape(()).to(torch.float32),
                        "exp_avg": self.flat.slice_of(p, self.exp_avg).view_as(p).clone(),
                        "exp_avg_sq": self.flat.slice_of(p, self.exp_avg_sq).view_as(p).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group.update(maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(len(self.flat.params))))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        for i, p in enumerate(self.flat.params):
            st = sd["state"].get(i)
            if st is not None:
                self.flat.slice_of(p, self.exp_avg).copy_(st["exp_avg"].reshape(-1))
                self.flat.slice_of(p, self.exp_avg_sq).copy_(st["exp_avg_sq"].reshape(-1))
                self.step_count.fill_(float(st["step"]))
        for k in ("lr", "betas", "eps"):
            if k in sd["param_groups"][0]:
                self.param_groups[0][k] = sd["param_groups"][0][k]


class DataParallelStepper:
    """The rank-symmetric step protocol (module docstring).  Subclasses provide the compute:
      optimistic_pass(batch) -> loss   forward + backward, failure words into self.pool (no host sync)
      checked_pass(batch)    -> loss   forward + backward with immediate checks; raises F.NotPSDError
      update(veto)                     optimiser update (veto: device int words or None)
    and own `self.grads` (FlatGrads) and `self.pool` (F.InfoPool)."""

    def exchange(self):
        """Steps 2 + 3: failure words first (MAX), then the gradient (SUM) - both by every rank, unconditionally."""
        if _world() > 1:
            dist.all_reduce(self.pool.words(), op=dist.ReduceOp.MAX)
            self.grads.allreduce_sum()

    def device_step(self, batch):
        """Steps 1-4: no host synchronisation inside; capturable in a CUDA graph."""
        self.pool.reset()
        loss = self.optimistic_pass(batch)
        self.exchange()
        self.update(self.pool.words())
        return loss

    def resolve(self, batch, loss):
        """Step 5, after device_step or a replay of it."""
        if not self.pool.any_failed():
            return loss
        return self.collective_checked_step(batch)

    def collective_checked_step(self, batch):
        self.grads.zero()
        err, loss = None, None
        early, self.grads.early_enabled = self.grads.early_enabled, False   # a rank may raise before the hook fires
        try:
            loss = self.checked_pass(batch)
        except F.NotPSDError as e:
            err = e
        finally:
            self.grads.early_enabled = early
        flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=self.grads.flat.device)
        if _world() > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            self.grads.zero()
            raise err if err is not None else F.NotPSDError("a Cholesky factorisation failed on another rank")
        self.grads.allreduce_sum()
        self.update(None)
        return loss


class ElboTrainer(DataParallelStepper):
    """VariationalCMP (+CMPLikelihood) or VariationalGP (+VBaggGaussianLikelihood) ELBO steps."""

    def __init__(self, model, likelihood, num_data, beta=1.0, lr=0.01, cuda_graph=False):
        self.model, self.likelihood = model, likelihood
        self.elbo = BagVariationalELBO(likelihood, model, num_data=num_data, beta=beta)
        self.params = list(model.parameters()) + list(likelihood.parameters())
        # q(u) and lazily drawn buffers must exist before the optimizer / flat buffers are laid out and broadcast
        model.variational_strategy.ensure_initialized()
        materialize_lazy_buffers(model)
        self.grads = FlatParams(self.params)
        sync_module_state([model, likelihood], self.grads.values)
        self.opt = FusedAdam(self.grads, lr=lr)
        self.grads.overlap_largest()
        self.pool = F.InfoPool(self.grads.flat.device)
        model.train()
        likelihood.train()
        # CUDA-graph mode: steps 1-4 of the protocol for a minibatch SHAPE are captured once (third call with that
        # shape) and replayed
        self.cuda_graph = cuda_graph
        self._graphs, self._graph_calls, self._capture_stream = {}, {}, None
        self.launches_replayed = 0   # libdcgp kernel launches executed through graph replays (host counters do not see them)

    def _fused_vbagg_loss(self, x, z, bags_sizes):
        """-ELBO of a VBagg minibatch through the two-call C composite (dcgp_vbagg_elbo_fwd / _bwd), or None when the
        model is outside its scope: additive stationary kernel terms only, zero or constant mean.  FP32 models keep the
        reference's FP64 K_ZZ island inside the composite.  DCGP_VBAGG_FUSED=0 keeps the layered path."""
        from .likelihoods import bag_offsets
        from .means import ConstantMean, ZeroMean
        vs = getattr(self.model, "variational_strategy", None)
        if vs is None or not hasattr(vs, "_prior_modules"):
            return None
        if os.environ.get("DCGP_VBAGG_FUSED", "1") == "0":
            return None
        kernel, mean = vs._prior_modules()
        x2 = x if x.dim() == 2 else x.unsqueeze(-1)
        kops = kernel.ops(x2)
        if not kops.is_flat or not isinstance(mean, (ConstantMean, ZeroMean)):
            return None
        vs.ensure_initialized()
        vd = vs._variational_distribution
        mc = mean.constant.to(x.dtype) if isinstance(mean, ConstantMean) else None
        elbo = F.vbagg_elbo(kops.meta, kops.hyp, vs.inducing_points.to(x.dtype), x2, bag_offsets(bags_sizes, x.device), z,
                            vd.variational_mean.to(x.dtype), vd.chol_variational_covar.to(x.dtype),
                            self.likelihood.noise.to(x.dtype), self.elbo.num_data, self.elbo.beta, mean_const=mc)
        return -elbo

    def _fused_cmp_loss(self, x, z, bags_values, extended_bags_values):
        """-ELBO of a VariationalCMP minibatch through the two-call C composite (dcgp_cmp_elbo_fwd / _bwd: the FP64 K_ZZ
        island, the Lambda factorisation and both solves, the bag-level Gaussian terms and the whole backward pass below
        the C ABI), or None when the model is outside its scope: stationary additive kernel terms only, zero or constant
        individuals mean, individuals noise.  DCGP_CMP_FUSED=0 keeps the layered path."""
        from .means import ConstantMean, ZeroMean
        if os.environ.get("DCGP_CMP_FUSED", "1") == "0":
            return None
        m = self.model
        if not (hasattr(m, "bag_kernel") and hasattr(m, "individuals_kernel")) or getattr(m, "noise_kernel", None) is None:
            return None
        if not getattr(self.likelihood, "use_individuals_noise", False):
            return None
        mean = m.individuals_mean
        x2 = x if x.dim() == 2 else x.unsqueeze(-1)
        y2 = bags_values if bags_values.dim() == 2 else bags_values.unsqueeze(-1)
        e2 = extended_bags_values if extended_bags_values.dim() == 2 else extended_bags_values.unsqueeze(-1)
        kk, kl = m.individuals_kernel.ops(x2), m.bag_kernel.ops(e2)
        if not (kk.is_flat and kl.is_flat) or not isinstance(mean, (ConstantMean, ZeroMean)):
            return None
        vs = m.variational_strategy
        vs.ensure_initialized()
        vd = vs._variational_distribution
        dt = x.dtype
        mc = mean.constant.to(dt) if isinstance(mean, ConstantMean) else None
        elbo = F.cmp_elbo(kk.meta, kk.hyp, kl.meta, kl.hyp, vs.inducing_points.to(dt), x2, e2, y2, z, vd.variational_mean.to(dt),
                          vd.chol_variational_covar.to(dt), self.likelihood.noise.to(dt), m.noise_kernel.outputscale.to(dt),
                          m.lbda, self.elbo.num_data, self.elbo.beta, mean_const=mc)
        return -elbo

    def loss(self, x, z, **kw):
        if isinstance(self.likelihood, CMPLikelihood) and set(kw) == {"bags_values", "extended_bags_values"}:
            try:
                fused = self._fused_cmp_loss(x, z, kw["bags_values"], kw["extended_bags_values"])
            except F.NotPSDError:
                fused = None      # checked pass after a failed factorisation: the layered path retries with jitter
            if fused is not None:
                return fused
        if isinstance(self.likelihood, VBaggGaussianLikelihood) and set(kw) == {"bags_sizes"}:
            try:
                fused = self._fused_vbagg_loss(x, z, kw["bags_sizes"])
            except F.NotPSDError:
                fused = None      # checked pass after a failed factorisation: the layered path retries with jitter
            if fused is not None:
                return fused
        q = self.model(x)
        if isinstance(self.likelihood, CMPLikelihood):
            kw = self.model.get_elbo_computation_parameters(bags_values=kw["bags_values"],
                                                            extended_bags_values=kw["extended_bags_values"])
        return -self.elbo(variational_dist_f=q, target=z, **kw)

    # ---- DataParallelStepper compute ---------------------------------------------------------------------
    def optimistic_pass(self, batch):
        with F.deferred_pd_checks(self.pool):
            loss = self.loss(batch["x"], batch["z"], **batch["kw"])
            loss.backward()
        return loss.detach()

    def checked_pass(self, batch):
        loss = self.loss(batch["x"], batch["z"], **batch["kw"])
        loss.backward()
        return loss.detach()

    def update(self, veto):
        # gradients are averaged over the ranks inside the update; the buffer is left zeroed for the next pass
        self.opt.step(veto=veto, grad_scale=1.0 / _world(), zero_grads=True)

    # ---- public step ----------------------------------------------------------------------------------------
    def step(self, x, z, **kw):
        """One ELBO step on device-resident tensors; returns the (device) loss."""
        if self.cuda_graph:
            return self._graph_step(x, z, kw)
        return self._eager_step(x, z, kw)

    def _eager_step(self, x, z, kw):
        batch = {"x": x, "z": z, "kw": kw}
        return self.resolve(batch, self.device_step(batch))

    # ---- CUDA-graph mode ---------------------------------------------------------------------------------
    def _graph_step(self, x, z, kw):
        """Everything of graph mode - eager warm-up steps, capture, replays - runs on ONE dedicated stream:
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

    def _graph_step_on_stream(self, x, z, kw):
        tensors = {"x": x, "z": z, **{k: v for k, v in kw.items() if torch.is_tensor(v)}}
        others = {k: v for k, v in kw.items() if not torch.is_tensor(v)}
        g = self.opt.param_groups[0]
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(tensors.items())) + \
            tuple((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in sorted(others.items())) + \
            (("lr", g["lr"]),)        # the learning rate is a launch constant of the captured update
        entry = self._graphs.get(key)
        if entry is None:
            n = self._graph_calls[key] = self._graph_calls.get(key, 0) + 1
            if n <= 2:   # eager warm-up: kernel attributes, look-ahead lanes, allocator
                return self._eager_step(x, z, kw)
            # a failed capture is an error, not a silent fall-back to eager steps under the "graph" label
            entry = self._graphs[key] = self._capture(tensors, others)
        static, graph, loss, launches = entry
        for k, v in tensors.items():
            if v is not static[k]:
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        self.launches_replayed += launches
        batch = {"x": static["x"], "z": static["z"], "kw": {**{k: v for k, v in static.items() if k not in ("x", "z")}, **others}}
        return self.resolve(batch, loss.clone())

    def _capture(self, tensors, others):
        """Called on the capture stream."""
        from . import _lib
        dev = tensors["x"].device
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        static = {k: v.clone() for k, v in tensors.items()}
        batch = {"x": static["x"], "z": static["z"], "kw": {**{k: v for k, v in static.items() if k not in ("x", "z")}, **others}}
        torch.cuda.synchronize(dev)
        lib = _lib.load()
        graph = torch.cuda.CUDAGraph()
        n0 = lib.dcgp_launch_count()
        with torch.cuda.graph(graph, stream=self._capture_stream):
            loss = self.device_step(batch)
        launches = int(lib.dcgp_launch_count() - n0)
        return static, graph, loss, launches

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
