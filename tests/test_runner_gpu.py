"""A miniature experiment written the way the reference's runners are (docopt usage block, YAML config, MODELS /
TRAINERS registries, `import gpytorch`, `from src.models import ...`, progress bars, running_logs.yaml and state.pt
dumps; cf. experiments/swiss_roll/run_experiment.py:34-75,186-206 and experiments/swiss_roll/models/
{exact_cmp,variational_cmp,vbagg}.py), launched UNCHANGED through `deconditional_downscaling_b200.runner` on the GPU.
The reference checkout itself does not exist on the GPU box, hence this stand-in experiment; the real
run_experiment.py is driven through the same launcher in tests/test_runner_cpu.py up to the first kernel call."""
import os
import subprocess
import sys

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUN_PY = '''
"""
Description : miniature bagged-regression experiment

Usage: run.py  [options] --cfg=<path_to_config> --o=<output_dir>

Options:
  --cfg=<path_to_config>           YAML configuration file.
  --o=<output_dir>                 Output directory.
  --device=<device_index>          GPU index [default: 0].
  --lr=<lr>                        Learning rate.
  --n_epochs=<n_epochs>            Number of training epochs
  --plot                           Outputs scatter plots.
"""
import os
import yaml
import logging
from docopt import docopt
import torch
import matplotlib.pyplot as plt
from models import build_model, train_model


def make_dataset(cfg):
    g = torch.Generator().manual_seed(cfg["seed"])
    sizes = torch.randint(cfg["min_bag"], cfg["max_bag"], (cfg["n_bags"],), generator=g).tolist()
    n = sum(sizes)
    x = torch.randn(n, 3, generator=g)
    f = torch.sin(x[:, 0]) + 0.5 * x[:, 1] * x[:, 2]
    centres = torch.linspace(-2, 2, cfg["n_bags"])
    y = centres.unsqueeze(-1)
    ext = torch.cat([c.repeat(s) for c, s in zip(centres, sizes)]).unsqueeze(-1)
    x[:, 0] = x[:, 0] + ext[:, 0]
    f = torch.sin(x[:, 0]) + 0.5 * x[:, 1] * x[:, 2]
    z = torch.stack([c.mean() for c in f.split(sizes)]) + 0.05 * torch.randn(cfg["n_bags"], generator=g)
    return sizes, x, ext, y, z, f


if __name__ == "__main__":
    args = docopt(__doc__)
    with open(args["--cfg"], "r") as fh:
        cfg = yaml.safe_load(fh)
    if args["--lr"]:
        cfg["training"]["lr"] = float(args["--lr"])
    if args["--n_epochs"]:
        cfg["training"]["n_epochs"] = int(args["--n_epochs"])
    logging.basicConfig(level=logging.INFO)
    os.makedirs(args["--o"], exist_ok=True)
    with open(os.path.join(args["--o"], "cfg.yaml"), "w") as fh:
        yaml.dump(cfg, fh)
    sizes, x, ext, y, z, f = make_dataset(cfg["dataset"])
    cfg["model"].update(individuals=x, extended_bags_values=ext, bags_values=y, aggregate_targets=z, bags_sizes=sizes)
    model = build_model(cfg["model"])
    logging.info(f"Initialized model {model}")
    cfg["training"].update(model=model, individuals=x, extended_bags_values=ext, bags_values=y, aggregate_targets=z,
                           bags_sizes=sizes, groundtruth_individuals=x, groundtruth_targets=f,
                           device_idx=args["--device"], dump_dir=args["--o"])
    train_model(cfg["training"])
'''

MODELS_PY = '''
import os
import yaml
import numpy as np
import torch
import gpytorch
from progress.bar import Bar
from src.models import ExactCMP, VariationalCMP, VariationalGP
from src.likelihoods import CMPLikelihood, VBaggGaussianLikelihood
from src.mlls import BagVariationalELBO
from src.kernels import RFFKernel
from src.utils import Registry

MODELS, TRAINERS = Registry(), Registry()


def build_model(cfg):
    return MODELS[cfg["name"]](**cfg)


def train_model(cfg):
    return TRAINERS[cfg["name"]](**cfg)


def _kernels():
    inv_softplus = lambda v, n: torch.log(torch.exp(v * torch.ones(n)) - 1)  # noqa: E731
    k = gpytorch.kernels.RBFKernel(ard_num_dims=3)
    k.initialize(raw_lengthscale=inv_softplus(1., 3))
    l = gpytorch.kernels.RBFKernel(ard_num_dims=1)
    l.initialize(raw_lengthscale=inv_softplus(1., 1))
    return gpytorch.kernels.ScaleKernel(k), gpytorch.kernels.ScaleKernel(l)


def _inducing(individuals, m, seed):
    g = torch.Generator().manual_seed(seed)
    return individuals[torch.randperm(len(individuals), generator=g)[:m]].clone().float()


@MODELS.register("exact_cmp")
def build_exact_cmp(individuals, extended_bags_values, bags_values, aggregate_targets, lbda, use_individuals_noise, **kwargs):
    k, l = _kernels()
    return ExactCMP(individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=k, bag_kernel=l,
                    train_individuals=individuals, train_bags=bags_values, train_aggregate_targets=aggregate_targets,
                    extended_train_bags=extended_bags_values, lbda=lbda, use_individuals_noise=use_individuals_noise,
                    likelihood=gpytorch.likelihoods.GaussianLikelihood())


@MODELS.register("variational_cmp")
def build_variational_cmp(n_inducing_points, lbda, individuals, use_individuals_noise, seed, **kwargs):
    k, l = _kernels()
    return VariationalCMP(individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=k, bag_kernel=l,
                          inducing_points=_inducing(individuals, n_inducing_points, seed), lbda=lbda,
                          use_individuals_noise=use_individuals_noise)


@MODELS.register("vbagg")
def build_vbagg(n_inducing_points, individuals, seed, **kwargs):
    k, _ = _kernels()
    return VariationalGP(inducing_points=_inducing(individuals, n_inducing_points, seed),
                         mean_module=gpytorch.means.ZeroMean(), covar_module=k)


def _predict(model, individuals):
    with torch.no_grad():
        return model.predict(individuals) if hasattr(model, "predict") else model(individuals)


def _epoch_logs(model, x, t, chunk_size):
    exact = isinstance(model, ExactCMP)   # the exact model predicts in training mode and clears its cache afterwards
    if not exact:
        model.eval()
    post = _predict(model, x)
    logs = {"mse": torch.pow(post.mean - t, 2).mean().item(), "mae": torch.abs(post.mean - t).mean().item()}
    try:
        nlls = []
        with gpytorch.settings.fast_computations(log_prob=False, covar_root_decomposition=False):
            for X, tt in zip(x.split(chunk_size), t.split(chunk_size)):
                nlls.append((-_predict(model, X).log_prob(tt).div(chunk_size)).item())
        logs["nll"] = float(np.mean(nlls))
    except gpytorch.utils.errors.NotPSDError:
        logs["nll"] = float("nan")
    if exact:
        model._clear_cache()
    else:
        model.train()
    return logs


def _dump(dump_dir, logs, model, optimizer, n_epochs):
    with open(os.path.join(dump_dir, "running_logs.yaml"), "w") as fh:
        yaml.dump({"epoch": logs}, fh)
    torch.save({"epoch": n_epochs, "state_dict": model.state_dict(), "optimizer": optimizer.state_dict()},
               os.path.join(dump_dir, "state.pt"))


@TRAINERS.register("exact_cmp")
def train_exact_cmp(model, lr, n_epochs, groundtruth_individuals, groundtruth_targets, device_idx, chunk_size, dump_dir, **kwargs):
    device = torch.device(f"cuda:{device_idx}") if torch.cuda.is_available() else torch.device("cpu")
    x, t = groundtruth_individuals.to(device), groundtruth_targets.to(device)
    model = model.to(device)
    model.train()
    model.likelihood.train()
    optimizer = torch.optim.Adam(params=model.parameters(), lr=lr)
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(model.likelihood, model)
    bar, logs = Bar("Epoch", max=n_epochs), dict()
    for epoch in range(n_epochs):
        optimizer.zero_grad()
        output = model(model.train_bags)
        loss = -mll(output, model.train_aggregate_targets)
        loss.backward()
        optimizer.step()
        model.update_cmp_estimate_parameters()
        bar.suffix = f"NLL {loss.item()}"
        bar.next()
        logs[epoch + 1] = dict(_epoch_logs(model, x, t, chunk_size), loss=loss.item())
    _dump(dump_dir, logs, model, optimizer, n_epochs)


@TRAINERS.register("variational_cmp")
def train_variational_cmp(model, individuals, extended_bags_values, bags_values, aggregate_targets, use_individuals_noise,
                          lr, n_epochs, beta, seed, groundtruth_individuals, groundtruth_targets, device_idx, chunk_size,
                          dump_dir, **kwargs):
    device = torch.device(f"cuda:{device_idx}") if torch.cuda.is_available() else torch.device("cpu")
    individuals, extended_bags_values = individuals.to(device), extended_bags_values.to(device)
    bags_values, aggregate_targets = bags_values.to(device), aggregate_targets.to(device)
    x, t = groundtruth_individuals.to(device), groundtruth_targets.to(device)
    likelihood = CMPLikelihood(use_individuals_noise=use_individuals_noise)
    model, likelihood = model.train().to(device), likelihood.train().to(device)
    optimizer = torch.optim.Adam(params=list(model.parameters()) + list(likelihood.parameters()), lr=lr)
    elbo = BagVariationalELBO(likelihood, model, num_data=len(aggregate_targets), beta=beta)
    bar, logs = Bar("Epoch", max=n_epochs), dict()
    torch.random.manual_seed(seed)
    for epoch in range(n_epochs):
        idx = torch.randperm(len(individuals))
        xe, ye = individuals[idx], extended_bags_values[idx]
        optimizer.zero_grad()
        q = model(xe)
        kw = model.get_elbo_computation_parameters(bags_values=bags_values, extended_bags_values=ye)
        loss = -elbo(variational_dist_f=q, target=aggregate_targets, **kw)
        loss.backward()
        optimizer.step()
        bar.suffix = f"ELBO {-loss.item()}"
        bar.next()
        logs[epoch + 1] = dict(_epoch_logs(model, x, t, chunk_size), loss=loss.item(),
                               aggregate_noise=likelihood.noise.detach().item(),
                               k_outputscale=model.individuals_kernel.outputscale.detach().item(),
                               k_lengthscale_x=model.individuals_kernel.base_kernel.lengthscale.detach()[0].tolist()[0])
    _dump(dump_dir, logs, model, optimizer, n_epochs)


@TRAINERS.register("vbagg")
def train_vbagg(model, individuals, aggregate_targets, bags_sizes, lr, n_epochs, beta, groundtruth_individuals,
                groundtruth_targets, device_idx, chunk_size, dump_dir, **kwargs):
    device = torch.device(f"cuda:{device_idx}") if torch.cuda.is_available() else torch.device("cpu")
    individuals, aggregate_targets = individuals.to(device), aggregate_targets.to(device)
    x, t = groundtruth_individuals.to(device), groundtruth_targets.to(device)
    likelihood = VBaggGaussianLikelihood()
    model, likelihood = model.train().to(device), likelihood.train().to(device)
    optimizer = torch.optim.Adam(params=list(model.parameters()) + list(likelihood.parameters()), lr=lr)
    elbo = BagVariationalELBO(likelihood, model, num_data=len(aggregate_targets), beta=beta)
    bar, logs = Bar("Epoch", max=n_epochs), dict()
    for epoch in range(n_epochs):
        optimizer.zero_grad()
        q = model(individuals)
        loss = -elbo(variational_dist_f=q, target=aggregate_targets, bags_sizes=bags_sizes)
        loss.backward()
        optimizer.step()
        bar.suffix = f"ELBO {-loss.item()}"
        bar.next()
        logs[epoch + 1] = dict(_epoch_logs(model, x, t, chunk_size), loss=loss.item())
    _dump(dump_dir, logs, model, optimizer, n_epochs)


@MODELS.register("downscaling_vcmp")
def build_downscaling_vcmp(individuals, n_inducing_points, lbda, use_individuals_noise, num_samples, **kwargs):
    # the downscaling experiment's kernel zoo: random-feature Matern / RBF sums on column subsets for the individuals,
    # exact Matern + RBF for the bags, inducing points on a regular stride through the individuals
    inv_softplus = lambda v, n: torch.log(torch.exp(v * torch.ones(n)) - 1)  # noqa: E731
    k_spatial = RFFKernel(nu=1.5, num_samples=num_samples, ard_num_dims=2, active_dims=[0, 1])
    k_spatial.initialize(raw_lengthscale=inv_softplus(1., 2))
    k_feat = gpytorch.kernels.RFFKernel(num_samples=num_samples, ard_num_dims=1, active_dims=[2])
    k_feat.initialize(raw_lengthscale=inv_softplus(1., 1))
    individuals_kernel = gpytorch.kernels.ScaleKernel(k_spatial) + gpytorch.kernels.ScaleKernel(k_feat)
    l_a = gpytorch.kernels.MaternKernel(nu=1.5, ard_num_dims=1, active_dims=[0])
    l_a.initialize(raw_lengthscale=inv_softplus(1., 1))
    l_b = gpytorch.kernels.RBFKernel(ard_num_dims=1, active_dims=[0])
    l_b.initialize(raw_lengthscale=inv_softplus(2., 1))
    bag_kernel = gpytorch.kernels.ScaleKernel(l_a) + gpytorch.kernels.ScaleKernel(l_b)
    n = individuals.size(0)
    step, offset = n // n_inducing_points, (n % n_inducing_points) // 2
    inducing_points = individuals[offset:n - offset:step].float()
    return VariationalCMP(individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=individuals_kernel,
                          bag_kernel=bag_kernel, inducing_points=inducing_points, lbda=lbda,
                          use_individuals_noise=use_individuals_noise)


@TRAINERS.register("downscaling_vcmp")
def train_downscaling_vcmp(model, individuals, extended_bags_values, bags_values, aggregate_targets, use_individuals_noise,
                           lr, n_epochs, beta, seed, batch_size, batch_size_cme, groundtruth_individuals, groundtruth_targets,
                           device_idx, dump_dir, **kwargs):
    device = torch.device(f"cuda:{device_idx}") if torch.cuda.is_available() else torch.device("cpu")
    individuals, extended_bags_values = individuals.to(device), extended_bags_values.to(device)
    bags_values, aggregate_targets = bags_values.to(device), aggregate_targets.to(device)
    grid, t = groundtruth_individuals.to(device), groundtruth_targets
    torch.random.manual_seed(seed)

    def batch_iterator(batch_size):
        def hr_iterator(batch_size):
            buffer = torch.ones(len(individuals)).to(device)
            while True:
                idx = buffer.multinomial(batch_size)
                yield individuals[idx], extended_bags_values[idx]
        sampler = hr_iterator(batch_size_cme)
        for idx in torch.randperm(len(aggregate_targets)).to(device).split(batch_size):
            x, extended_y = next(sampler)
            yield x, bags_values[idx], extended_y, aggregate_targets[idx]

    likelihood = CMPLikelihood(use_individuals_noise=use_individuals_noise)
    model, likelihood = model.train().to(device), likelihood.train().to(device)
    optimizer = torch.optim.Adam(params=list(model.parameters()) + list(likelihood.parameters()), lr=lr)
    elbo = BagVariationalELBO(likelihood, model, num_data=len(aggregate_targets), beta=beta)
    mean_shift, std_scale = t.mean().item(), t.std().item()
    epoch_bar, logs = Bar("Epoch", max=n_epochs), dict()
    for epoch in range(n_epochs):
        batch_bar = Bar("Batch", max=len(aggregate_targets) // batch_size)
        epoch_loss, n_batches = 0, 0
        for x, y, extended_y, z in batch_iterator(batch_size):
            optimizer.zero_grad()
            q = model(x)
            kw = model.get_elbo_computation_parameters(bags_values=y, extended_bags_values=extended_y)
            loss = -elbo(variational_dist_f=q, target=z, **kw)
            loss.backward()
            optimizer.step()
            epoch_loss += loss.item()
            n_batches += 1
            batch_bar.suffix = f"Running ELBO {-loss.item()}"
            batch_bar.next()
        # posterior over the whole grid, rescaled and moved to the host as the experiment's predictor does
        model.eval()
        with torch.no_grad():
            post = model(grid)
        mean = mean_shift + std_scale * post.mean.cpu()
        lazy_cov = (std_scale ** 2) * post.lazy_covariance_matrix.cpu()
        post = gpytorch.distributions.MultivariateNormal(mean=mean, covariance_matrix=lazy_cov)
        model.train()
        lower, upper = post.confidence_region()
        k_spatial, k_feat = model.individuals_kernel.kernels
        l_a, l_b = model.bag_kernel.kernels
        logs[epoch + 1] = {"loss": epoch_loss / n_batches, "mse": torch.pow((post.mean - mean_shift) / std_scale - (t - mean_shift) / std_scale, 2).mean().item(),
                           "ci_width": (upper - lower).mean().item(), "aggregate_noise": likelihood.noise.detach().item(),
                           "k_spatial_outputscale": k_spatial.outputscale.detach().item(),
                           "k_lengthscale_lat": k_spatial.base_kernel.lengthscale[0].detach().tolist()[0],
                           "k_feat_outputscale": k_feat.outputscale.detach().item(),
                           "l_lengthscale": l_a.base_kernel.lengthscale[0].detach().tolist()[0],
                           "l_b_outputscale": l_b.outputscale.detach().item(),
                           "indiv_noise": model.noise_kernel.outputscale.detach().item() if model.noise_kernel else None}
        epoch_bar.next()
    _dump(dump_dir, logs, model, optimizer, n_epochs)
'''

CFGS = {
    "exact_cmp": {"model": {"name": "exact_cmp", "lbda": 0.01, "use_individuals_noise": False},
                  "training": {"name": "exact_cmp", "lr": 0.05, "n_epochs": 12, "chunk_size": 200}},
    "variational_cmp": {"model": {"name": "variational_cmp", "n_inducing_points": 48, "lbda": 0.001,
                                  "use_individuals_noise": True, "seed": 3},
                        "training": {"name": "variational_cmp", "lr": 0.05, "n_epochs": 12, "beta": 1.0, "seed": 3,
                                     "use_individuals_noise": True, "chunk_size": 200}},
    "vbagg": {"model": {"name": "vbagg", "n_inducing_points": 48, "seed": 3},
              "training": {"name": "vbagg", "lr": 0.05, "n_epochs": 12, "beta": 1.0, "chunk_size": 200}},
    # the downscaling experiment's shape in miniature: stochastic minibatches of bags + a multinomial CME sample,
    # random-feature additive kernels, individuals noise (experiments/downscaling/models/variational_cmp.py:15-228)
    "downscaling_vcmp": {"model": {"name": "downscaling_vcmp", "n_inducing_points": 40, "lbda": 0.001,
                                   "use_individuals_noise": True, "num_samples": 64},
                         "training": {"name": "downscaling_vcmp", "lr": 0.05, "n_epochs": 12, "beta": 1.0, "seed": 3,
                                      "use_individuals_noise": True, "batch_size": 8, "batch_size_cme": 200}},
}

EXPECTED_KEYS = {
    "exact_cmp": ["likelihood.noise_covar.raw_noise", "individuals_kernel.base_kernel.raw_lengthscale",
                  "bag_kernel.raw_outputscale", "mean_module.bag_kernel.base_kernel.raw_lengthscale"],
    "variational_cmp": ["variational_strategy.inducing_points",
                        "variational_strategy._variational_distribution.variational_mean",
                        "variational_strategy._variational_distribution.chol_variational_covar",
                        "individuals_kernel.base_kernel.raw_lengthscale", "bag_kernel.raw_outputscale", "noise_kernel.raw_outputscale"],
    "vbagg": ["variational_strategy.inducing_points", "variational_strategy._variational_distribution.chol_variational_covar",
              "covar_module.base_kernel.raw_lengthscale"],
    "downscaling_vcmp": ["variational_strategy.inducing_points", "individuals_kernel.kernels.0.base_kernel.raw_lengthscale",
                         "individuals_kernel.kernels.1.raw_outputscale", "bag_kernel.kernels.0.base_kernel.raw_lengthscale",
                         "bag_kernel.kernels.1.raw_outputscale", "noise_kernel.raw_outputscale"],
}


@pytest.fixture(scope="module")
def experiment(tmp_path_factory):
    root = tmp_path_factory.mktemp("mini_experiment")
    (root / "models").mkdir()
    (root / "config").mkdir()
    (root / "run.py").write_text(RUN_PY)
    (root / "models" / "__init__.py").write_text(MODELS_PY)
    for name, cfg in CFGS.items():
        full = {"dataset": {"seed": 5, "n_bags": 16, "min_bag": 20, "max_bag": 41}, **cfg}
        (root / "config" / f"{name}.yaml").write_text(yaml.dump(full))
    return root


@pytest.mark.parametrize("name", ["exact_cmp", "variational_cmp", "vbagg", "downscaling_vcmp"])
def test_unchanged_style_runner_trains_and_dumps_reference_format(experiment, tmp_path, name):
    out = tmp_path / name
    r = subprocess.run([sys.executable, "-m", "deconditional_downscaling_b200.runner", str(experiment / "run.py"),
                        f"--cfg={experiment / 'config' / (name + '.yaml')}", f"--o={out}", "--device=0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env={**os.environ, "PYTHONPATH": ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")})
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Epoch 12/12" in r.stderr                                   # the progress-bar stand-in was driven
    assert yaml.safe_load((out / "cfg.yaml").read_text())["model"]["name"] == name
    logs = yaml.safe_load((out / "running_logs.yaml").read_text())["epoch"]
    assert sorted(logs) == list(range(1, 13))
    losses = [logs[e]["loss"] for e in sorted(logs)]
    assert all(map(lambda v: v == v and abs(v) < 1e6, losses))        # finite
    assert min(losses[-3:]) < losses[0]                                # 12 epochs of Adam reduce the objective
    assert logs[12]["mse"] == logs[12]["mse"]
    if name == "downscaling_vcmp":
        assert logs[12]["ci_width"] > 0 and logs[12]["indiv_noise"] > 0
    else:
        assert logs[12]["nll"] == logs[12]["nll"]
    state = torch.load(out / "state.pt", weights_only=True)
    assert set(state) == {"epoch", "state_dict", "optimizer"} and state["epoch"] == 12
    for key in EXPECTED_KEYS[name]:
        assert key in state["state_dict"], (key, sorted(state["state_dict"]))
