"""CPU test: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/dcgp.h declares (no compute call is made without a GPU)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dcgp.h")).read()
    return sorted(set(re.findall(r"DCGP_API\s+[\w ]+?\s+(dcgp_\w+)\(", text)))


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "dcgp.h"\nint main(void){dcgp_kernel_spec s; s.n_components = 1; return (int)sizeof(s) == 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_library_builds_and_exports_every_declared_symbol():
    from deconditional_downscaling_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dcgp.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    # identity queries are host-only and safe without a GPU
    assert lib.dcgp_arch() == 1000
    assert lib.dcgp_version() >= 100


def test_sass_is_sm100a_and_uses_tensor_pipe():
    from deconditional_downscaling_b200 import build
    path = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "DMMA" in sass, "FP64 contraction must be on the DMMA tensor path"


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from deconditional_downscaling_b200 import ops, _lib
    spec = ops.FlatSpec(["rbf"], [[0]], [torch.ones(1)], [torch.ones(1)])
    with pytest.raises(_lib.DcgpError):
        ops.gram(spec, torch.zeros(4, 1))


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """The ctypes mirrors in _lib.py (KernelSpec, CmpProblem, CmpGrads, VbaggProblem, VbaggGrads) have the size and the
    field offsets of the C structs in include/dcgp.h - checked by compiling a C program that prints them."""
    from deconditional_downscaling_b200 import _lib
    pairs = [("dcgp_kernel_spec", _lib.KernelSpec), ("dcgp_component", _lib.Component), ("dcgp_cmp_problem", _lib.CmpProblem),
             ("dcgp_cmp_grads", _lib.CmpGrads), ("dcgp_vbagg_problem", _lib.VbaggProblem), ("dcgp_vbagg_grads", _lib.VbaggGrads)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dcgp.h"', 'int main(void){']
    for cname, ct in pairs:
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in pairs:
        assert int(out[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_workspace_queries_are_host_only_and_monotone():
    """Workspace sizes of the composites are plain host arithmetic (no GPU needed): positive, growing with every
    dimension, zero for degenerate shapes, and large enough for the buffers the headers promise."""
    from deconditional_downscaling_b200 import _lib
    lib = _lib.load()
    for dt, es in ((_lib.F32, 4), (_lib.F64, 8)):
        base = lib.dcgp_cmp_elbo_workspace(2048, 256, 128, 3, dt)
        assert base >= (2 * 2048 * 2048 + 5 * 2048 * 256 + 4 * 128 * 2048) * es
        assert lib.dcgp_cmp_elbo_workspace(4096, 256, 128, 3, dt) > base
        assert lib.dcgp_cmp_elbo_workspace(2048, 512, 128, 3, dt) > base
        assert lib.dcgp_cmp_elbo_workspace(2048, 256, 256, 3, dt) > base
        assert lib.dcgp_cmp_elbo_workspace(0, 256, 128, 3, dt) == 0
        vb = lib.dcgp_vbagg_elbo_workspace(256, 100, 3, dt)
        assert vb >= (4 * 256 * 256 + 3 * 256 * 100) * es
        assert lib.dcgp_vbagg_elbo_workspace(512, 100, 3, dt) > vb
        assert lib.dcgp_vbagg_elbo_workspace(256, 0, 3, dt) == 0
        assert lib.dcgp_gram_bwd_lowrank_workspace(300, 200, dt) == 300 * 200 * es
    # the FP32 composites carry the FP64 island: double copies of Z and of M x M matrices
    assert lib.dcgp_vbagg_elbo_workspace(256, 100, 8, _lib.F32) > lib.dcgp_vbagg_elbo_workspace(256, 100, 3, _lib.F32)
