"""CPU simulation of candidate operand-precision schemes of the conv GEMMs (diagnostic; uses the
oracle, so it lives under tools/ and is never imported by the product).

  fp16      : A_hi*B_hi                                         (1 MMA unit)
  fp16w2    : A_hi*(B_hi+B_lo)                                  (2 units)
  fp16x3    : A_hi*B_hi + A_hi*B_lo + A_lo*B_hi                 (3 units)
  f16f8     : A_hi*B_hi [fp16] + (A_hi8*B_lo8 + A_lo8*B_hi8) [e4m3, K-concatenated, 2x rate]  (2 units)

  i8x2      : int8 Ozaki split, two 8-bit slices per operand, hi*hi + hi*lo + lo*hi [kind::i8, 2x rate]  (1.5 units)
  i8x3      : three slices per operand, the six products of weight >= 2^-16                              (3 units)
              (A scaled per tensor -- the K vector of a 3x3 conv spans neighbouring pixels, so a per-pixel scale
              cannot be factored out -- B per output channel)

All accumulation in fp64 here: only the operand rounding is simulated."""
import sys, os, time
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.resnet50_unet import OracleNet, BN_EPS
from sbb_textline_detection_b200 import synth
from sbb_textline_detection_b200.detector import synthetic_weights

F8 = torch.float8_e4m3fn

def q16(t): return t.to(torch.float16).to(t.dtype)
def q8(t, fmt=F8, lim=448.0): return t.clamp(-lim, lim).to(torch.float32).to(fmt).to(t.dtype)

class SimNet(OracleNet):
    def __init__(self, w, nc, mode, sa=0.25):
        super().__init__(w, nc, dtype=torch.float64, quant=None)
        self.mode, self.sa = mode, sa
    def _split_act(self, x):
        hi = q16(x); lo = x - hi
        return hi, lo
    def _conv(self, x, name, stride=1, pad=0, bn=None, relu=False, add=None):
        k, b = self._k[name], self.w[name + "/bias"]
        if bn is not None:
            s = self.w[bn + "/gamma"] / torch.sqrt(self.w[bn + "/var"] + BN_EPS)
            k = k * s[:, None, None, None]
            b = (b - self.w[bn + "/mean"]) * s + self.w[bn + "/beta"]
        k = k.to(torch.float32).to(torch.float64)  # host packs from fp32-rounded folded weights? keep fp64->hi/lo
        conv = lambda a, w: F.conv2d(a, w, None, stride=stride, padding=pad)
        m = self.mode
        if isinstance(m, dict):  # per-layer scheme: {"default": ..., "<layer>": ...}
            m = m.get(name, m["default"])
        bh = q16(k); bl = q16(k - bh)
        is_input = name in ("conv1",)  # image enters as exact hi+lo fp16 (packed), all modes but fp16
        ah, al = self._split_act(x)
        if m == "exact":
            y = conv(x, k)
        elif m == "fp16":
            y = conv(ah, bh)
        elif m == "fp16w2":
            y = conv(ah, bh + bl)
        elif m == "fp16x3":
            y = conv(ah, bh) + conv(ah, bl) + conv(q16(al), bh)
        elif m in ("i8x2", "i8x3"):
            ns = 2 if m == "i8x2" else 3
            full = float(2 ** (8 * ns - 1) - 1)
            sa = x.abs().max().clamp_min(1e-30) / full
            sb = k.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30) / full
            ai, bi = torch.round(x / sa), torch.round(k / sb)

            def slices(v):  # signed 8-bit digits, most significant first: v = sum d_i * 256^(ns-1-i)
                out, rest = [], v
                for i in range(ns - 1, 0, -1):
                    lo = rest - 256.0 * torch.round(rest / 256.0)
                    out.insert(0, lo)
                    rest = (rest - lo) / 256.0
                out.insert(0, rest)
                return out
            A, B = slices(ai), slices(bi)
            y = 0
            for i in range(ns):
                for j in range(ns):
                    if i + j <= ns - 1:  # drop the products below 2^-(8*ns) of the leading one
                        y = y + conv(A[i], B[j]) * (256.0 ** (2 * (ns - 1) - i - j))
            y = y * sa * sb.reshape(1, -1, 1, 1)
        elif m.startswith("f16f8"):
            if is_input:
                y = conv(ah, bh) + conv(ah, bl) + conv(q16(al), bh)
            else:
                sa = self.sa
                sb = 2.0 ** np.floor(np.log2(256.0 / float(k.abs().max())))
                X = 4096.0
                ah8 = q8(ah * sa); al8 = q8(al * sa * X)
                bh8 = q8(bh * sb); bl8 = q8((k - bh) * sb * X)
                if "A" in m: ah8 = ah * sa; bl8 = (k - bh) * sb * X     # cross1 exact
                if "B" in m: al8 = al * sa * X; bh8 = bh * sb           # cross2 exact
                if "a" in m[5:]: ah8 = ah * sa
                if "b" in m[5:]: bh8 = bh * sb
                if "l" in m[5:]: al8 = al * sa * X; bl8 = (k - bh) * sb * X
                y = conv(ah, bh) + (conv(ah8, bl8) + conv(al8, bh8)) / (sa * sb * X)
        y = y + b[None, :, None, None]
        if add is not None:
            if isinstance(m, str) and m.startswith("f16f8") and "R" not in m:
                rh, rl = self._split_act(add)
                add = rh + q8(rl * self.sa * 4096.0) / (self.sa * 4096.0)
            elif m in ("fp16", "fp16w2"):
                add = q16(add)
            y = y + add
        if relu:
            y = F.relu(y)
        return y

MIXED = {  # is fp32-grade arithmetic needed in EVERY layer?  cheaper schemes in the last decoder blocks only
    "dec5:w2": {"default": "fp16x3", "dec5": "fp16w2"},
    "dec5:fp16": {"default": "fp16x3", "dec5": "fp16"},
    "dec4-5:w2": {"default": "fp16x3", "dec5": "fp16w2", "dec4": "fp16w2"},
    "dec3-5:w2": {"default": "fp16x3", "dec5": "fp16w2", "dec4": "fp16w2", "dec3": "fp16w2"},
    "dec1-5:w2": {"default": "fp16x3", "dec5": "fp16w2", "dec4": "fp16w2", "dec3": "fp16w2", "dec2": "fp16w2", "dec1": "fp16w2"},
}


def main():
    tile = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp16", "fp16w2", "fp16x3", "f16f8"] + list(MIXED)
    w, nc = synthetic_weights("textline")
    page = synth.document_page(1000, 1000, seed=3)
    x = np.stack([page[:tile, :tile], page[300:300 + tile, 400:400 + tile]]).astype(np.float32) / np.float32(255)
    with torch.no_grad():
        z64 = OracleNet(w, nc, dtype=torch.float64).logits(x)
        z32 = OracleNet(w, nc).logits(x).to(torch.float64)
        print(f"tile {tile}: fp32 oracle vs fp64: max {float((z32 - z64).abs().max()):.3e}")
        for m in modes:
            t = time.time()
            z = SimNet(w, nc, MIXED.get(m, m)).logits(x)
            e64 = (z - z64).abs(); e32 = (z - z32).abs()
            print(f"{m:10s} vs fp64: max {float(e64.max()):.3e} rms {float(e64.pow(2).mean().sqrt()):.3e} | vs fp32 oracle: max {float(e32.max()):.3e}   ({time.time()-t:.0f}s)", flush=True)

if __name__ == "__main__":
    main()
