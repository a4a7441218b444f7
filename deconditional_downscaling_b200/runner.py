"""Launcher for the reference's UNCHANGED experiment scripts.

    python -m deconditional_downscaling_b200.runner <reference>/experiments/swiss_roll/run_experiment.py \
        --cfg=<reference>/experiments/swiss_roll/config/exact_cmp.yaml --o=out/

does what `python experiments/swiss_roll/run_experiment.py --cfg=... --o=...` does in a checkout of the reference
(README.md:24-33 there), with the model classes of this package behind the `gpytorch` / `src` names
(`compat.install()`), and with light stand-ins for the three third-party modules the scripts import but this image
does not ship:

* `docopt` (experiments/*/run_experiment.py:25,201-203): the scripts only use `--name=<value>` options, flags and
  `[default: x]` annotations of the "Options:" block, which is what `docopt()` below parses;
* `progress.bar.Bar` (experiments/swiss_roll/models/variational_cmp.py:96-132 and the other trainers):
  `Bar(message, max=)`, `.suffix`, `.next()`, `.finish()`, written to stderr once per `next()`;
* `matplotlib.pyplot` (run_experiment.py:27; core/visualization.py): imported unconditionally, used only under
  `--plot`; the stand-in imports fine and raises on first use.

A stand-in is installed only when the real module is not importable.  The script then runs with
`runpy.run_path(..., run_name="__main__")`, its own directory first on `sys.path` (the scripts do
`import core.generation`, `from models import ...`).  Compute goes through libdcgp; there is no CPU fallback, so
without a CUDA device the first kernel call raises.
"""
import importlib.util
import os
import re
import runpy
import sys
import types


# --------------------------------------------------------------------------------------------------- docopt
class DocoptExit(SystemExit):
    def __init__(self, message, usage=""):
        super().__init__((message + "\n" + usage).strip())


_OPTION_LINE = re.compile(r"^\s*(--[\w-]+)(?:[ =](<[^>]+>|[A-Z][A-Z_]*))?(?:\s{2,}|\s*$)(.*)$")
_DEFAULT = re.compile(r"\[default: (.*?)\]", re.IGNORECASE)


def _parse_options_block(doc):
    """{'--name': (takes_value, default)} from the lines that follow 'Options:'."""
    spec, inside = {}, False
    for line in doc.splitlines():
        if re.match(r"^\s*options:\s*$", line, re.IGNORECASE):
            inside = True
            continue
        if not inside:
            continue
        m = _OPTION_LINE.match(line)
        if not m:
            continue
        name, value, desc = m.group(1), m.group(2), m.group(3)
        d = _DEFAULT.search(desc or "")
        spec[name] = (value is not None, d.group(1) if (d and value is not None) else None)
    return spec


def _usage_section(doc):
    m = re.search(r"^\s*usage:.*?(?=^\s*$)", doc, re.IGNORECASE | re.MULTILINE | re.DOTALL)
    return m.group(0).strip() if m else ""


def docopt(doc, argv=None, help=True, version=None, options_first=False):
    """The subset of docopt the reference's runners rely on: long options only, `--name=<v>` / `--name <v>` /
    flags, defaults from `[default: x]`, options named in the usage pattern outside `[options]` are required.
    Returns {'--name': value-string | None | bool}."""
    argv = list(sys.argv[1:] if argv is None else argv)
    spec = _parse_options_block(doc)
    usage = _usage_section(doc)
    # options written in the usage line itself (e.g. --cfg=<path_to_config>) are required and take a value
    required = []
    for name, value in re.findall(r"(--[\w-]+)(?:=(<[^>]+>))?", re.sub(r"\[[^\]]*\]", " ", usage)):
        required.append(name)
        spec.setdefault(name, (bool(value), None))
    out = {name: (default if takes else False) for name, (takes, default) in spec.items()}
    i = 0
    while i < len(argv):
        tok = argv[i]
        i += 1
        if help and tok in ("-h", "--help"):
            print(doc.strip("\n"))
            raise SystemExit(0)
        if version is not None and tok == "--version":
            print(version)
            raise SystemExit(0)
        if not tok.startswith("--"):
            raise DocoptExit(f"unexpected argument {tok!r}", usage)
        name, eq, val = tok.partition("=")
        if name not in spec:
            # unambiguous prefixes are accepted, as in docopt
            hits = [n for n in spec if n.startswith(name)]
            if len(hits) != 1:
                raise DocoptExit(f"{name} is not a unique prefix of a known option" if hits else f"unknown option {name}", usage)
            name = hits[0]
        takes, _ = spec[name]
        if takes:
            if not eq:
                if i >= len(argv):
                    raise DocoptExit(f"{name} requires an argument", usage)
                val = argv[i]
                i += 1
            out[name] = val
        else:
            if eq:
                raise DocoptExit(f"{name} must not have an argument", usage)
            out[name] = True
    missing = [n for n in required if out.get(n) in (None, False)]
    if missing:
        raise DocoptExit("missing required option(s): " + ", ".join(missing), usage)
    return out


# --------------------------------------------------------------------------------------------------- progress.bar
class Bar:
    """progress.bar.Bar as the trainers drive it: a counter with a message, a maximum and a free-text suffix."""

    def __init__(self, message="", max=100, file=None, **kwargs):
        self.message, self.max, self.index, self.suffix = message, max, 0, ""
        self.file = sys.stderr if file is None else file
        self.quiet = os.environ.get("DCGP_QUIET_PROGRESS", "") not in ("", "0")

    def _write(self, end="\n"):
        if not self.quiet and self.file is not None:
            print(f"{self.message} {self.index}/{self.max} {self.suffix}".rstrip(), file=self.file, end=end, flush=True)

    def next(self, n=1):
        self.index += n
        self._write()

    def goto(self, index):
        self.index = index
        self._write()

    def finish(self):
        pass

    def iter(self, it):
        for x in it:
            yield x
            self.next()
        self.finish()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.finish()


# --------------------------------------------------------------------------------------------------- matplotlib
class _UnavailableModule(types.ModuleType):
    """Imports fine, fails on first attribute use (plotting is outside the hot path)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        raise ImportError(f"{self.__name__}.{name}: matplotlib is not installed in this environment; "
                          "run without --plot or install matplotlib")


def _importable(name):
    if name in sys.modules:
        return True
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install_standins():
    """Register docopt / progress / matplotlib stand-ins for the modules that are NOT importable.  Returns their names."""
    done = []
    if not _importable("docopt"):
        m = types.ModuleType("docopt")
        m.docopt, m.DocoptExit = docopt, DocoptExit
        sys.modules["docopt"] = m
        done.append("docopt")
    if not _importable("progress"):
        p, b = types.ModuleType("progress"), types.ModuleType("progress.bar")
        b.Bar = b.IncrementalBar = b.ChargingBar = Bar
        p.bar, p.__path__ = b, []
        sys.modules["progress"], sys.modules["progress.bar"] = p, b
        done.append("progress")
    if not _importable("matplotlib"):
        root = _UnavailableModule("matplotlib")
        root.__path__ = []
        for sub in ("pyplot", "cm", "colors", "patches", "lines", "gridspec", "ticker"):
            child = _UnavailableModule(f"matplotlib.{sub}")
            root.__dict__[sub] = child
            sys.modules[f"matplotlib.{sub}"] = child
        sys.modules["matplotlib"] = root
        done.append("matplotlib")
    return done


def run(script, argv=(), force=False):
    """Execute `script` (a path to one of the reference's run_experiment.py files, or any script written against
    the reference's `gpytorch` / `src` API) as __main__ with `argv` as its command line."""
    from . import compat
    script = os.path.abspath(script)
    if not os.path.isfile(script):
        raise FileNotFoundError(script)
    compat.install(force=force)
    install_standins()
    saved_argv, saved_path = sys.argv, list(sys.path)
    sys.argv = [script] + list(argv)
    sys.path.insert(0, os.path.dirname(script))
    try:
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv, sys.path[:] = saved_argv, saved_path


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    run(argv[0], argv[1:])
    return 0


if __name__ == "__main__":
    sys.exit(main())
