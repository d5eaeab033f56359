"""Layer inventory of the ResNet50-U-Net that the reference loads from its .h5 files
(main.py:58-60, 216-223; architecture from qurator-spk/sbb_pixelwise_segmentation, see
SURVEY.md Appendix A).  Order == the order in which the packed weight blob stores the layers and
the order in which csrc/sbb_net.cu consumes them."""
from __future__ import annotations

from dataclasses import dataclass

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon

STAGES = (
    (2, "abc", (64, 64, 256), 1),
    (3, "abcd", (128, 128, 512), 2),
    (4, "abcdef", (256, 256, 1024), 2),
    (5, "abc", (512, 512, 2048), 2),
)


@dataclass(frozen=True)
class ConvSpec:
    name: str
    bn: str
    kh: int
    kw: int
    cin: int
    cout: int


def conv_specs(n_classes: int):
    specs = [ConvSpec("conv1", "bn_conv1", 7, 7, 3, 64)]
    cin = 64
    for stage, blocks, (f1, f2, f3), _ in STAGES:
        for b in blocks:
            base, bnb = f"res{stage}{b}_branch", f"bn{stage}{b}_branch"
            specs.append(ConvSpec(base + "2a", bnb + "2a", 1, 1, cin, f1))
            specs.append(ConvSpec(base + "2b", bnb + "2b", 3, 3, f1, f2))
            specs.append(ConvSpec(base + "2c", bnb + "2c", 1, 1, f2, f3))
            if b == "a":
                specs.append(ConvSpec(base + "1", bnb + "1", 1, 1, cin, f3))
            cin = f3
    specs += [
        ConvSpec("dec_v5", "bn_dec_v5", 1, 1, 2048, 512),
        ConvSpec("dec_v4", "bn_dec_v4", 1, 1, 1024, 512),
        ConvSpec("dec1", "bn_dec1", 3, 3, 1024, 512),
        ConvSpec("dec2", "bn_dec2", 3, 3, 1024, 256),
        ConvSpec("dec3", "bn_dec3", 3, 3, 512, 128),
        ConvSpec("dec4", "bn_dec4", 3, 3, 192, 64),
        ConvSpec("dec5", "bn_dec5", 3, 3, 67, 32),
        ConvSpec("cls", "bn_cls", 1, 1, 32, n_classes),
    ]
    return specs


def tile_geometry(tile_h: int, tile_w: int):
    """Spatial size of every feature level for one tile: dict level -> (h, w).
    level 0 = input, 1 = conv1/f1, 2 = stage2 (after 'valid' maxpool), 3..5 = stages 3..5."""
    assert tile_h % 32 == 0 and tile_w % 32 == 0, "tile size must be a multiple of 32"
    g = {0: (tile_h, tile_w)}
    g[1] = ((tile_h + 6 - 7) // 2 + 1, (tile_w + 6 - 7) // 2 + 1)
    g[2] = ((g[1][0] - 3) // 2 + 1, (g[1][1] - 3) // 2 + 1)
    for s in (3, 4, 5):
        g[s] = ((g[s - 1][0] - 1) // 2 + 1, (g[s - 1][1] - 1) // 2 + 1)
    return g


def conv_flops_per_tile(tile_h: int, tile_w: int, n_classes: int):
    """ALGORITHMIC work per tile = 2*MACs of every Conv2D.
    Returns (total, encoder, decoder) in FLOP.  448x448, C=2 -> 87.85e9 (30.75 + 57.10)."""
    g = tile_geometry(tile_h, tile_w)
    px = {k: v[0] * v[1] for k, v in g.items()}
    enc = 2 * px[1] * 147 * 64
    cin = 64
    for stage, blocks, (f1, f2, f3), _ in STAGES:
        for b in blocks:
            enc += 2 * px[stage] * (cin * f1 + 9 * f1 * f2 + f2 * f3)
            if b == "a":
                enc += 2 * px[stage] * cin * f3
            cin = f3
    dec = 2 * px[5] * 2048 * 512 + 2 * px[4] * 1024 * 512
    dec += 2 * px[4] * 9 * 1024 * 512
    dec += 2 * (4 * px[4]) * 9 * 1024 * 256
    dec += 2 * (16 * px[4]) * 9 * 512 * 128
    dec += 2 * (64 * px[4]) * 9 * 192 * 64
    dec += 2 * (256 * px[4]) * 9 * 67 * 32
    dec += 2 * (256 * px[4]) * 32 * n_classes
    return enc + dec, enc, dec
