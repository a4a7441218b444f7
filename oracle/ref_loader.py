"""TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference package
(/root/reference/src) with `gpytorch` aliased to oracle.gpytorch_min (the
restated gpytorch==1.4.0 subset).  /root/reference exists only in the build
container: callers must skip when `reference_available()` is False.  Used by
tests/golden/make_golden.py to generate the committed golden vectors."""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIM = "oracle.gpytorch_min"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def load_reference_src():
    if not reference_available():
        raise RuntimeError("reference sources are not mounted (build container only)")
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    import oracle.gpytorch_min as shim  # noqa: F401
    sys.modules["gpytorch"] = shim
    for name, mod in list(sys.modules.items()):
        if name.startswith(_SHIM + "."):
            sys.modules["gpytorch." + name[len(_SHIM) + 1:]] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src  # the reference's package
    if not os.path.abspath(src.__file__).startswith(REFERENCE_ROOT):
        raise RuntimeError(f"`src` resolved to {src.__file__}, not the reference")
    return src
