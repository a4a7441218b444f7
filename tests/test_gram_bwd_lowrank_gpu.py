"""dcgp_gram_bwd_lowrank: hyper-parameter adjoints of a Gram matrix for an upstream gradient given as a low-rank product
G = alpha P Q^T (the form in which the adjoints of Lambda and K_xx arrive in the ELBO step) against the dense route
(G formed by a product, then dcgp_gram_bwd) and against torch autograd in FP64."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dense_k(x1, x2, kinds, dims, ls, os_):
    out = 0
    for kind, dm, l, o in zip(kinds, dims, ls, os_):
        a, b = x1[:, dm] / l, x2[:, dm] / l
        r2 = (a[:, None, :] - b[None, :, :]).pow(2).sum(-1)
        if kind == "rbf":
            out = out + o * torch.exp(-0.5 * r2)
        else:
            r = r2.clamp_min(1e-30).sqrt()
            out = out + o * (1 + 3 ** 0.5 * r) * torch.exp(-(3 ** 0.5) * r)
    return out


CASES = [("rbf5", ["rbf"], [[0, 1, 2, 3, 4]]), ("add", ["matern32", "rbf"], [[0, 1], [2]]), ("rbf3", ["rbf"], [[0, 1, 2]])]


@pytest.mark.parametrize("name,kinds,dims", CASES)
@pytest.mark.parametrize("n1,n2,k,same", [(1024, 1024, 256, True), (700, 1300, 96, False), (4096, 4096, 1024, True), (130, 257, 33, False)])
def test_lowrank_adjoint_matches_dense_adjoint_and_autograd(name, kinds, dims, n1, n2, k, same):
    from deconditional_downscaling_b200 import ops
    g = torch.Generator().manual_seed(n1 + k)
    d = max(max(dm) for dm in dims) + 1
    x1 = torch.randn(n1, d, generator=g, dtype=torch.float64)
    x2 = x1 if same else torch.randn(n2, d, generator=g, dtype=torch.float64)
    n2 = x2.shape[0]
    P = torch.randn(n1, k, generator=g, dtype=torch.float64) / k ** 0.5
    Q = torch.randn(n2, k, generator=g, dtype=torch.float64)
    ls = [0.8 + torch.rand(len(dm), generator=g, dtype=torch.float64) for dm in dims]
    os_ = [torch.tensor(0.5 + 0.3 * i, dtype=torch.float64) for i in range(len(dims))]
    alpha = -0.7
    # FP64 autograd reference
    ls64 = [l.clone().requires_grad_(True) for l in ls]
    os64 = [o.clone().requires_grad_(True) for o in os_]
    G = alpha * P @ Q.t()
    (_dense_k(x1, x2, kinds, dims, ls64, os64) * G).sum().backward()
    scale = (_dense_k(x1, x2, kinds, dims, ls, os_) * G.abs()).sum().item()      # magnitude of the signed sum
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 3e-4)):
        t = lambda v: v.to(device=DEV, dtype=dtype)
        spec = ops.FlatSpec(kinds, dims, [t(l) for l in ls], [t(o) for o in os_])
        Pd, Qd = t(P), t(Q)
        if dtype == torch.float32:
            Pd, Qd = ops.round_tf32(Pd), ops.round_tf32(Qd)
        for fused in ((False, True) if dtype == torch.float32 else (False,)):
            dls, dos = ops.gram_bwd_lowrank(spec, t(x1), None if same else t(x2), Pd, Qd, alpha=alpha, fused=fused)
            for c in range(len(dims)):
                nd = len(dims[c])
                assert (dls[c, :nd].double().cpu() - ls64[c].grad).abs().max().item() <= tol * scale, (name, dtype, fused, "ls", c)
                assert abs(dos[c].item() - os64[c].grad.item()) <= tol * scale, (name, dtype, fused, "os", c)
        if dtype == torch.float32:
            # the dense route on the same rounded operands (single-pass product, then the adjoint kernel)
            Gd = alpha * (Pd.double() @ Qd.double().t()).float()
            dls2, dos2, _ = ops.gram_bwd(spec, t(x1), None if same else t(x2), Gd)
            assert (dls - dls2).abs().max().item() <= 3e-4 * scale
            assert (dos - dos2).abs().max().item() <= 3e-4 * scale
