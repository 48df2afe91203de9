"""Layer-by-layer comparison of the engine's workspace against the oracle's
intermediate activations (test infrastructure; needs a GPU).

    python tests/gpu_debug.py [--simt] [--shape 1,1,8,16,8] [--downs 2] [--ngf 16]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import unet_oracle as O  # noqa: E402


def buffer_producers(cfg):
    """[(buffer index, tap module index, channel slice)] mirroring build_program()
    in anatomix_b200/csrc/engine.cu."""
    nd, g = cfg["num_downs"], cfg["ngf"]
    width = [g << i for i in range(nd + 1)]
    out, nxt = [], nd                       # buffers 0..nd-1 are the concat buffers
    out.append((nxt, 2, None)); nxt += 1    # stem output after its activation
    for i in range(nd):
        e = 3 + 7 * i
        out.append((nxt, e + 2, None)); nxt += 1          # first conv of the level
        out.append((nxt, e + 6, None)); nxt += 1          # pooled skip
    b = 3 + 7 * nd
    out.append((nxt, b + 2, None)); nxt += 1
    out.append((nxt, b + 5, None)); nxt += 1
    for j in range(nd):
        l = nd - 1 - j
        d = b + 6 + 7 * j
        out.append((l, d, None))                           # concat buffer == tap at the Upsample index
        out.append((nxt, d + 3, None)); nxt += 1
        out.append((nxt, d + 6, None)); nxt += 1
    return out


def to_padded_planar(t):
    """[N,C,D,H,W] fp32 -> reflect-padded planar [N, C/8, D+2, H+2, W+2, 8]."""
    p = F.pad(t, (1, 1, 1, 1, 1, 1), mode="reflect")
    n, c, d, h, w = p.shape
    return p.view(n, c // 8, 8, d, h, w).permute(0, 1, 3, 4, 5, 2).contiguous()


def read_padded(eng, ws, off, shape, dtype):
    """[N, G, D+2, H+2, W+2, 8] view of a padded planar buffer (rows carry lead / tail voxels)."""
    n, g, dp, hp, wp, _ = shape
    lead, pitch = eng.row_layout(wp - 2)
    t = ws[off:off + n * g * dp * hp * pitch * 16].view(dtype).float().cpu().view(n, g, dp, hp, pitch, 8)
    return t[:, :, :, :, lead:lead + wp].contiguous()


def compare(eng, cfg, state, x, emulate=True):
    n, _, d, h, w = x.shape
    prods = buffer_producers(cfg)
    taps_ids = sorted({p[1] for p in prods})
    _, taps = O.unet_forward(cfg, state, x, layers=taps_ids, engine_rounding=emulate)
    tap = dict(zip(taps_ids, taps))
    ws = eng.workspace(n, d, h, w)
    table = eng.buffer_table(n, d, h, w)
    lines = []
    for buf, idx, _ in prods:
        off, nb, lvl, grp = table[buf]
        want = to_padded_planar(tap[idx])
        got = read_padded(eng, ws, off, want.shape, O.engine_storage_dtype(cfg))
        inner = (slice(None), slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
        den = want.norm().clamp_min(1e-20)
        rel_all = ((got - want).norm() / den).item()
        rel_in = ((got[inner] - want[inner]).norm() / want[inner].norm().clamp_min(1e-20)).item()
        bad = (~torch.isfinite(got)).sum().item()
        lines.append(f"buf {buf:2d} <- module {idx:2d} level {lvl} ch {grp * 8:4d}: "
                     f"rel-L2 interior {rel_in:.3e}  with shell {rel_all:.3e}  nonfinite {bad}")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--simt", action="store_true")
    ap.add_argument("--shape", default="1,1,8,16,8")
    ap.add_argument("--downs", type=int, default=2)
    ap.add_argument("--ngf", type=int, default=16)
    ap.add_argument("--seed", type=int, default=11)
    a = ap.parse_args()
    from anatomix_b200 import _lib
    from anatomix_b200.engine import Engine
    cfg = dict(dimension=3, input_nc=1, output_nc=16, num_downs=a.downs, ngf=a.ngf)
    full = dict(O.DEFAULTS); full.update(cfg)
    state = O.random_state(cfg, seed=a.seed)
    shape = tuple(int(s) for s in a.shape.split(","))
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(3))
    eng = Engine(cfg, "cuda:0", flags=(_lib.FLAG_FORCE_SIMT if a.simt else 0) | _lib.FLAG_NO_WS_REUSE | _lib.FLAG_NO_UPCONV)
    eng.load_state(state)
    y = eng.forward(x.cuda())
    torch.cuda.synchronize()
    print(compare(eng, full, state, x))
    # detail: second encoder conv of level 0 (module 8, first 16 channels of concat buffer 0) and its pool
    n, _, d, h, w = x.shape
    _, taps = O.unet_forward(cfg, state, x, layers=[5, 8, 9], engine_rounding=True)
    ws = eng.workspace(n, d, h, w)
    table = eng.buffer_table(n, d, h, w)
    def planar(buf, groups_total, dims):
        off, nb, lvl, grp = table[buf]
        dd, hh, ww = dims
        return read_padded(eng, ws, off, (n, groups_total, dd + 2, hh + 2, ww + 2, 8), torch.bfloat16)
    cat0 = planar(0, 3 * a.ngf // 8, (d, h, w))[:, :a.ngf // 8]
    got8 = cat0.permute(0, 1, 5, 2, 3, 4).reshape(n, a.ngf, d + 2, h + 2, w + 2)[:, :, 1:-1, 1:-1, 1:-1]
    for name, got, want in (("conv8", got8, taps[1]),):
        diff = (got - want).abs()
        bad = diff > 0
        print(f"{name}: {int(bad.sum())} of {bad.numel()} elements differ; max abs {diff.max().item():.4g}")
        idx = bad.nonzero()[:12]
        for i in idx:
            i = tuple(i.tolist())
            print("   at", i, "got", got[i].item(), "want", want[i].item())
    for emu in (False, True):
        want = O.unet_forward(cfg, state, x, engine_rounding=emu)
        rel = ((y.cpu() - want).norm() / want.norm()).item()
        print(f"output vs oracle (engine_rounding={emu}): rel-L2 {rel:.3e}")


if __name__ == "__main__":
    main()
