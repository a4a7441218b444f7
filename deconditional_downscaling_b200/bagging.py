"""Data-side bag layout on the device: the grid -> bags reshaping of the downscaling
experiment and the bag bookkeeping the fused kernels consume.  Pure index arithmetic on tensors that already live in HBM -
one strided view / one copy per call instead of Python loops over blocks - no floating-point work, hence no libdcgp
kernel; everything here works on whatever device the inputs are on.

* `make_bagged_dataset` - experiments/downscaling/core/preprocessing.py:186-216 (same argument order, same four outputs,
  same element order; the reference's split loops become one view + permute, and `extended_bags` is an expand view
  unless `materialize=True`);
* `bags_to_offsets` / `offsets_to_bag_ids` - `bags_sizes` lists (experiments/swiss_roll/core/generation.py:60-98) as the
  int64 `offsets[n_bags + 1]` device array of the segmented kernels (dcgp_gram_bagsum, dcgp_vbagg_ell) and the row -> bag map;
* `extend_bags_values` - `extended_bags_values` of generation.py:90-97: every individual carries its bag's value.
"""
from typing import Sequence, Tuple

import torch


def make_bagged_dataset(covariates_grid: torch.Tensor, bags_grid: torch.Tensor, target_grid: torch.Tensor,
                        block_size: Tuple[int, int], materialize: bool = False):
    """(H, W, d) fine grid, (h, w, r) coarse grid, (h, w) targets -> covariates_blocks (n_bags, bag_size, d),
    bags_blocks (n_bags, r), extended_bags (n_bags, bag_size, r), targets_blocks (n_bags,).

    Block (a, b) holds fine rows [a * block_size[1], (a + 1) * block_size[1]) and fine columns
    [b * block_size[0], (b + 1) * block_size[0]) - the reference splits dim 0 by `block_width` and dim 1 by
    `block_height` (preprocessing.py:204-205; its blocks are square, 17 x 17) - flattened row-major, blocks ordered with
    the fine-row block index outermost.  As in the reference, ragged edges are not supported: H and W must be multiples
    of the block extents and n_bags must equal h * w."""
    block_height, block_width = block_size
    H, W, d = covariates_grid.shape
    rows, cols = block_width, block_height          # extents along dim 0 / dim 1, see the docstring
    if H % rows or W % cols:
        raise ValueError(f"grid {H} x {W} is not a multiple of the block extents {rows} x {cols}")
    nr, nc = H // rows, W // cols
    bag_size = rows * cols
    covariates_blocks = covariates_grid.reshape(nr, rows, nc, cols, d).permute(0, 2, 1, 3, 4).reshape(nr * nc, bag_size, d)
    bags_blocks = bags_grid.reshape(-1, bags_grid.size(-1))
    if bags_blocks.size(0) != nr * nc:
        raise ValueError(f"{bags_blocks.size(0)} coarse cells for {nr * nc} blocks")
    extended_bags = bags_blocks.unsqueeze(1).expand(-1, bag_size, -1)
    if materialize:
        extended_bags = extended_bags.contiguous()
    targets_blocks = target_grid.reshape(-1)
    return covariates_blocks, bags_blocks, extended_bags, targets_blocks


def bags_to_offsets(bags_sizes: Sequence[int], device=None) -> torch.Tensor:
    """[n_0, n_1, ...] -> int64 offsets [0, n_0, n_0 + n_1, ...] on `device` (bags are contiguous row ranges)."""
    sizes = torch.as_tensor(list(bags_sizes), dtype=torch.int64)
    if sizes.numel() and int(sizes.min()) < 0:
        raise ValueError("negative bag size")
    offsets = torch.zeros(sizes.numel() + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(sizes, 0)
    return offsets.to(device) if device is not None else offsets


def offsets_to_bag_ids(offsets: torch.Tensor) -> torch.Tensor:
    """Row -> bag index (int64, length offsets[-1]) on the device of `offsets`; empty bags own no row."""
    sizes = offsets[1:] - offsets[:-1]
    return torch.repeat_interleave(torch.arange(sizes.numel(), device=offsets.device), sizes)


def extend_bags_values(bags_values: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """(n_bags, r) or (n_bags,) bag values -> one row per individual (generation.py:90-97)."""
    return bags_values[offsets_to_bag_ids(offsets)]
