"""Host-side dispatch rules of flaxdiff_b200.ops that decide which kernel arrangement runs (no GPU needed)."""
import torch

from flaxdiff_b200 import ops


def _t(*shape):
    return torch.empty(shape, dtype=torch.bfloat16, device="meta")


def test_groupnorm_backward_fusion_follows_the_transposed_engine():
    """The first pass of the GroupNorm backward is fused into the data gradient's epilogue exactly where that
    convolution runs on the transposed engine: Cin a multiple of 64 but not of 256, <= 512,
    16-aligned images of >= 128 pixels, Cout a multiple of 64."""
    ok = [(64, 64, 64, 64), (32, 32, 128, 128), (64, 64, 320, 64), (16, 16, 192, 256), (16, 16, 384, 256)]
    for h, w, cin, cout in ok:
        assert ops._gn_fuse_default(_t(2, h, w, cin), _t(2, h, w, cout)), (h, w, cin, cout)
    no = [(16, 16, 256, 256),      # Cin multiple of 256: pixels-as-M engine
          (16, 16, 512, 512),
          (8, 8, 128, 128),        # < 128 pixels per image (and not 16-aligned)
          (16, 8, 192, 128),       # width not a multiple of 16
          (32, 32, 576, 128),      # Cin > 512
          (32, 32, 128, 96)]       # Cout not a multiple of 64
    for h, w, cin, cout in no:
        assert not ops._gn_fuse_default(_t(2, h, w, cin), _t(2, h, w, cout)), (h, w, cin, cout)


def test_colstats_constants():
    assert ops.COLSTATS_SLOTS >= 1 and ops.COLSTATS_MIN_PIXELS == 128 and ops.GN_FUSE_MIN_PIXELS == 128
