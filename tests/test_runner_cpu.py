"""`runner.py`: the launcher for the reference's unchanged experiment scripts -
its docopt / progress / matplotlib stand-ins and the way it hosts a script.  No compute here; the GPU side is
tests/test_runner_gpu.py."""
import io
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from deconditional_downscaling_b200 import runner

# the option block of a runner written like experiments/*/run_experiment.py:1-23 (own wording)
DOC = """
Description : toy experiment

Usage: run.py  [options] --cfg=<path_to_config> --o=<output_dir>

Options:
  --cfg=<path_to_config>           YAML configuration file.
  --o=<output_dir>                 Output directory.
  --device=<device_index>          GPU index [default: 0].
  --unmatched                        Matched or unmatched datasets.
  --lr=<lr>                        Learning rate.
  --n_epochs=<n_epochs>            Number of epochs
  --plot                           Outputs plots.
"""


def test_docopt_standin_options_defaults_and_flags():
    a = runner.docopt(DOC, argv=["--cfg=c.yaml", "--o=out", "--lr=0.1", "--plot"])
    assert a == {"--cfg": "c.yaml", "--o": "out", "--device": "0", "--unmatched": False, "--lr": "0.1",
                 "--n_epochs": None, "--plot": True}
    # separated values and unambiguous prefixes, as docopt accepts them
    b = runner.docopt(DOC, argv=["--cfg", "c.yaml", "--o", "out", "--unm", "--dev=3"])
    assert b["--cfg"] == "c.yaml" and b["--unmatched"] is True and b["--device"] == "3" and b["--plot"] is False


@pytest.mark.parametrize("argv", [["--o=out"], ["--cfg=c.yaml", "--o=out", "--nope"], ["--cfg=c.yaml", "--o=out", "stray"],
                                  ["--cfg=c.yaml", "--o=out", "--plot=1"], ["--cfg=c.yaml", "--o"]])
def test_docopt_standin_rejects_bad_command_lines(argv):
    with pytest.raises(SystemExit) as e:
        runner.docopt(DOC, argv=argv)
    assert e.value.code not in (0, None)


def test_docopt_help_exits_zero(capsys):
    with pytest.raises(SystemExit) as e:
        runner.docopt(DOC, argv=["--help"])
    assert e.value.code == 0 and "Usage:" in capsys.readouterr().out


def test_progress_bar_standin():
    out = io.StringIO()
    bar = runner.Bar("Epoch", max=3, file=out)
    bar.quiet = False
    for i in range(3):
        bar.suffix = f"ELBO {i}"
        bar.next()
    bar.finish()
    lines = out.getvalue().splitlines()
    assert bar.index == 3 and lines[-1] == "Epoch 3/3 ELBO 2" and len(lines) == 3


SCRIPT = '''
"""
Usage: run.py [options] --o=<output_dir>

Options:
  --o=<output_dir>     Output directory.
  --n=<n>              A number [default: 7].
"""
import os, sys
from docopt import docopt
from progress.bar import Bar
import matplotlib.pyplot as plt          # imported unconditionally by the runners, used only under --plot
import gpytorch
from gpytorch.kernels import RBFKernel, ScaleKernel
from src.models import ExactCMP, VariationalCMP
from src.likelihoods import CMPLikelihood, VBaggGaussianLikelihood
from src.utils import Registry
import helper                            # a sibling module: the script's directory is on sys.path

if __name__ == "__main__":
    args = docopt(__doc__)
    os.makedirs(args["--o"], exist_ok=True)
    k = ScaleKernel(RBFKernel(ard_num_dims=2))
    with open(os.path.join(args["--o"], "ok.txt"), "w") as f:
        f.write(f"{args['--n']} {helper.VALUE} {type(k).__module__} {gpytorch.__version__} {sys.argv[1:]}")
'''


def test_run_hosts_a_script_with_the_reference_namespaces(tmp_path):
    (tmp_path / "run.py").write_text(SCRIPT)
    (tmp_path / "helper.py").write_text("VALUE = 11\n")
    out = tmp_path / "out"
    # in a fresh interpreter: the launcher registers modules process-wide
    r = subprocess.run([sys.executable, "-m", "deconditional_downscaling_b200.runner", str(tmp_path / "run.py"), f"--o={out}"],
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    n, value, mod, version, argv = (out / "ok.txt").read_text().split(" ", 4)
    assert (n, value) == ("7", "11") and mod.startswith("deconditional_downscaling_b200") and version.endswith("+dcgp")
    assert argv == str([f"--o={out}"])


def test_matplotlib_standin_fails_only_on_use():
    if runner._importable("matplotlib") and not isinstance(sys.modules.get("matplotlib"), runner._UnavailableModule):
        pytest.skip("a real matplotlib is installed")
    saved = {k: v for k, v in sys.modules.items() if k == "matplotlib" or k.startswith("matplotlib.")}
    try:
        runner.install_standins()
        import matplotlib.pyplot as plt
        with pytest.raises(ImportError, match="matplotlib is not installed"):
            plt.savefig("x.png")
    finally:
        for k in [k for k in sys.modules if k == "matplotlib" or k.startswith("matplotlib.")]:
            del sys.modules[k]
        sys.modules.update(saved)


REF = "/root/reference/experiments/swiss_roll"


@pytest.mark.skipif(not os.path.isdir(REF) or torch.cuda.is_available(),
                    reason="needs a checkout of the reference and a host WITHOUT a GPU")
@pytest.mark.parametrize("cfg,extra", [("exact_cmp", []), ("variational_cmp", ["--unmatched"])])
def test_unmodified_reference_runner_reaches_the_first_kernel_call_and_fails_loudly(tmp_path, cfg, extra):
    """The reference's own run_experiment.py, unedited: option parsing, YAML, data generation, model and trainer
    construction all run on this package; with no CUDA device the first kernel call raises (no CPU fallback)."""
    r = subprocess.run([sys.executable, "-m", "deconditional_downscaling_b200.runner", f"{REF}/run_experiment.py",
                        f"--cfg={REF}/config/{cfg}.yaml", f"--o={tmp_path}", "--n_epochs=1", *extra],
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "there is no CPU fallback" in r.stderr, r.stderr[-2000:]
    assert f"{REF}/models/{cfg}.py" in r.stderr          # ... from inside the reference's trainer
    assert (tmp_path / "cfg.yaml").exists()              # the runner got as far as dumping its configuration
