"""Row-block sharded Gram products across GPUs: for out = k(x1, x2) @ W, rank g owns
rows [g N / G, (g + 1) N / G) of x1, assembles only its (N / G) x n panel from locally generated Gram
tiles, and ONE all-gather exchanges the panels (N x n values: 86 MB in FP64 at N = 32768, n = 328).
Used for K W, k(x*, x) (Lambda^{-1} B) at prediction time and Lambda assembly; the exact Cholesky
itself stays on one GPU (north_star)."""
from typing import Callable, Optional

import torch
import torch.distributed as dist


def row_block(n: int, rank: int, world: int):
    """Rows owned by `rank`: contiguous, sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_rows_matmul(panel_fn: Callable[[int, int], torch.Tensor], n_rows: int, n_cols: int, dtype, device,
                        group=None) -> torch.Tensor:
    """Generic driver: `panel_fn(lo, hi)` returns this rank's (hi - lo) x n_cols panel; panels are
    all-gathered into the full n_rows x n_cols result on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return panel_fn(0, n_rows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = row_block(n_rows, rank, world)
    mine = panel_fn(lo, hi).contiguous()
    sizes = [row_block(n_rows, r, world) for r in range(world)]
    maxr = max(h - l for l, h in sizes)
    # equal-sized exchange buffers (all_gather needs them); ragged tails are trimmed afterwards
    send = torch.zeros(maxr, n_cols, dtype=dtype, device=device)
    send[: hi - lo] = mine
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([recv[r][: h - l] for r, (l, h) in enumerate(sizes)], dim=0)


def kernel_matmul_panel(kops, x1, x2, Wm, lo, hi, diag_const=0.0):
    """Rows [lo, hi) of (k(x1, x2) [+ diagonal terms when x2 is None]) @ Wm.  A row block of a self-product is a
    CROSS product for the Gram kernel (x1[lo:hi] against all of x1), which knows no diagonal: the delta-kernel scale
    (ScaleKernel(DeltaKernel), src/models/cmp.py:61-62) and the nugget are added here on the block's own rows, so
    the result does not depend on the number of ranks."""
    if x2 is None and lo == 0 and hi == x1.size(0):
        return kops.matmul(x1, None, Wm, diag_const=diag_const)
    out = kops.matmul(x1[lo:hi], x1 if x2 is None else x2, Wm)
    if x2 is None:
        dd = kops._delta_scale()
        if dd is not None or diag_const:
            out = out + ((dd if dd is not None else 0.0) + diag_const) * Wm[lo:hi]
    return out


def sharded_kernel_matmul(kernel, x1: torch.Tensor, x2: Optional[torch.Tensor], W: torch.Tensor, group=None, diag_const=0.0):
    """(k(x1, x2) [+ delta terms + diag_const I when x2 is None]) @ W with x1's rows sharded over the ranks.  Every
    rank holds x1, x2 and W (they are O(N d) / O(N n)); only Gram tiles are distributed."""
    x1 = x1 if x1.dim() == 2 else x1.unsqueeze(-1)
    x2f = None if x2 is None else (x2 if x2.dim() == 2 else x2.unsqueeze(-1))
    kops = kernel.ops(x1)
    Wm = W if W.dim() == 2 else W.unsqueeze(-1)
    out = sharded_rows_matmul(lambda lo, hi: kernel_matmul_panel(kops, x1, x2f, Wm, lo, hi, diag_const), x1.size(0),
                              Wm.size(1), x1.dtype, x1.device, group)
    return out.squeeze(-1) if W.dim() == 1 else out
