"""Data-side bag layout helpers (`bagging.py`) against a loop restatement of
experiments/downscaling/core/preprocessing.py:186-216 and of the bag bookkeeping of
experiments/swiss_roll/core/generation.py:60-98.  Index arithmetic only, so it is checked on the host."""
import pytest
import torch

from deconditional_downscaling_b200 import bagging


def loop_bagged_dataset(covariates_grid, bags_grid, target_grid, block_size):
    """The reference's construction, restated: split dim 0 by block_width, each piece along dim 1 by block_height."""
    block_height, block_width = block_size
    bag_size, ndim = block_height * block_width, covariates_grid.size(-1)
    blocks = []
    for r0 in range(0, covariates_grid.size(0), block_width):
        strip = covariates_grid[r0:r0 + block_width]
        for c0 in range(0, covariates_grid.size(1), block_height):
            blocks.append(strip[:, c0:c0 + block_height].reshape(bag_size, ndim))
    covariates_blocks = torch.stack(blocks)
    bags_blocks = bags_grid.reshape(-1, bags_grid.size(-1))
    extended = torch.stack([b.repeat(bag_size, 1) for b in bags_blocks])
    return covariates_blocks, bags_blocks, extended, target_grid.reshape(-1)


@pytest.mark.parametrize("h,w,block", [(4, 6, (17, 17)), (3, 5, (2, 3)), (5, 2, (4, 1)), (1, 1, (3, 3))])
def test_make_bagged_dataset_matches_loop_construction(h, w, block):
    bh, bw = block
    g = torch.Generator().manual_seed(h * 100 + w)
    H, W = h * bw, w * bh                       # dim 0 is cut by block_width, dim 1 by block_height
    cov = torch.randn(H, W, 5, generator=g)
    bags = torch.randn(h, w, 3, generator=g)
    tgt = torch.randn(h, w, generator=g)
    got = bagging.make_bagged_dataset(cov, bags, tgt, block)
    ref = loop_bagged_dataset(cov, bags, tgt, block)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and torch.equal(a, b)
    assert got[2].stride(1) == 0                                   # expand view: no bag_size-fold copy
    assert bagging.make_bagged_dataset(cov, bags, tgt, block, materialize=True)[2].is_contiguous()


def test_make_bagged_dataset_rejects_ragged_grids():
    with pytest.raises(ValueError):
        bagging.make_bagged_dataset(torch.zeros(10, 9, 2), torch.zeros(3, 3, 1), torch.zeros(3, 3), (3, 3))
    with pytest.raises(ValueError):
        bagging.make_bagged_dataset(torch.zeros(9, 9, 2), torch.zeros(2, 3, 1), torch.zeros(2, 3), (3, 3))


def test_offsets_bag_ids_and_extended_values():
    sizes = [3, 0, 5, 1]
    off = bagging.bags_to_offsets(sizes)
    assert off.dtype == torch.int64 and off.tolist() == [0, 3, 3, 8, 9]
    ids = bagging.offsets_to_bag_ids(off)
    assert ids.tolist() == [0, 0, 0, 2, 2, 2, 2, 2, 3]
    y = torch.tensor([[1., 10.], [2., 20.], [3., 30.], [4., 40.]])
    ext = bagging.extend_bags_values(y, off)
    ref = torch.cat([y[i].repeat(n, 1) for i, n in enumerate(sizes)])
    assert torch.equal(ext, ref)
    assert bagging.bags_to_offsets([]).tolist() == [0]
    with pytest.raises(ValueError):
        bagging.bags_to_offsets([2, -1])
