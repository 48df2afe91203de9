"""CPU ORACLE (test infrastructure, never shipped, never on the product path).

A functional restatement of ``anatomix.model.network.Unet.forward`` (reference
``anatomix/model/network.py:467-548``) for 3-D inputs, written against
``torch.nn.functional`` on the CPU.  The reference keeps no arithmetic of its own:
every op is a ``torch.nn`` layer (torch==2.13.0 pinned in the reference's
``requirements.txt:10``; torch 2.11.0 CPU in this image, same operator
semantics), so the restatement calls the same ATen operators at the same call
sites and re-derives the flat layer order from the constructor logic
(network.py:309-465) independently of the product's `anatomix_b200.topology`.

Pinning: the reference ships no tests or golden vectors, so
the pins are outputs of the *unmodified reference module* imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py`` and
committed under ``tests/golden/``; ``tests/test_oracle.py`` checks this file (and
the plain-C restatement ``unet_ref.c``) against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

``engine_rounding=True`` reproduces where the B200 engine rounds to its storage
type (packed weights, stored activations; bf16 for BatchNorm networks, fp16 for
InstanceNorm networks whose activations are bounded by the normalisation) while
accumulating in fp32, which gives the tight parity gate of SURVEY.md section 8(c).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DEFAULTS = dict(ngf=24, norm="batch", final_act="none", activation="relu",
                pad_type="reflect", doubleconv=True, residual_connection=False,
                pooling="Max", interp="nearest", use_skip_connection=True,
                norm_eps=1e-5)


def layer_program(input_nc, output_nc, num_downs, ngf, norm, activation,
                  final_act, doubleconv, use_skip_connection):
    """List of (op, index, cin, cout) in Sequential order + skip index lists.

    Follows the constructor top to bottom: stem (network.py:309-326), encoder
    loop (:334-369), bottleneck (:372-400), decoder loop (:403-445), final conv
    (:450-464)."""
    prog, enc_idx, dec_idx = [], [], []
    has_n, has_a = norm != "none", activation != "none"

    def block(ci, co):
        prog.append(("conv", len(prog), ci, co))
        if has_n:
            prog.append(("norm", len(prog), co, co))
        if has_a:
            prog.append(("act", len(prog), co, co))

    block(input_nc, ngf)
    ch = ngf
    for i in range(num_downs):
        mult = 1 if i == 0 else 2
        block(ch, ch * mult)
        if doubleconv:
            block(ch * mult, ch * mult)
        enc_idx.append(len(prog) - 1)
        prog.append(("pool", len(prog), ch * mult, ch * mult))
        ch *= mult
    block(ch, ch * 2)
    if doubleconv:
        block(ch * 2, ch * 2)
    mult = 2 ** num_downs
    for i in range(num_downs):
        dec_idx.append(len(prog))
        prog.append(("up", len(prog), ngf * mult, ngf * mult))
        m = mult + mult // 2 if use_skip_connection else mult
        block(ngf * m, ngf * (mult // 2))
        if doubleconv:
            block(ngf * (mult // 2), ngf * (mult // 2))
        mult //= 2
    prog.append(("conv", len(prog), ngf * mult, output_nc))
    if final_act != "none":
        prog.append(("final_act", len(prog), output_nc, output_nc))
    return prog, enc_idx, dec_idx


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def engine_storage_dtype(cfg):
    """Storage type the engine picks for a configuration."""
    c = dict(DEFAULTS)
    c.update(cfg)
    return torch.float16 if c["norm"] == "instance" else torch.bfloat16


def _activate(x, kind):
    if kind == "relu":
        return F.relu(x)                       # network.py:188-189
    if kind == "lrelu":
        return F.leaky_relu(x, 0.3)            # network.py:190-191 (slope 0.3)
    if kind == "elu":
        return F.elu(x)
    if kind == "selu":
        return F.selu(x)
    if kind == "tanh":
        return torch.tanh(x)
    raise ValueError(kind)


def conv3d_reflect(x, w, b):
    """``nn.Conv3d(k=3, s=1, padding='same', padding_mode='reflect')``
    (network.py:310-318): reflect-pad one voxel per side, valid correlation."""
    return F.conv3d(F.pad(x, (1, 1, 1, 1, 1, 1), mode="reflect"), w, b)


def _lowtap(par, k):
    """Low-resolution tap (0..2) hit by high-resolution tap k (0..2) at output parity par:
    floor((par + k - 1) / 2) + 1."""
    return (par + k - 1) // 2 + 1


def upconv_lowres(low, w_up, shift, rnd):
    """Engine emulation of conv(nearest_up2(low)) at the decoder's last level: per output parity
    (a, b, c) a 3x3x3 conv of the replicate-padded LOW-resolution tensor with parity-summed weights
    (summed in fp32, then rounded to the storage type), shift added, result rounded (it is stored as
    16-bit partial sums).  Reflect padding of the upsampled tensor == replicate padding of `low`."""
    n, _, d, h, w = low.shape
    co = w_up.shape[0]
    lp = F.pad(low, (1, 1, 1, 1, 1, 1), mode="replicate")
    out = torch.empty(n, co, 2 * d, 2 * h, 2 * w)
    for a in range(2):
        for b in range(2):
            for c in range(2):
                wp = torch.zeros_like(w_up)
                for kz in range(3):
                    for ky in range(3):
                        for kx in range(3):
                            wp[:, :, _lowtap(a, kz), _lowtap(b, ky), _lowtap(c, kx)] += w_up[:, :, kz, ky, kx]
                out[:, :, a::2, b::2, c::2] = F.conv3d(lp, rnd(wp)) + shift.view(1, -1, 1, 1, 1)
    return rnd(out)


@torch.no_grad()
def unet_forward(cfg, state, x, layers=(), training=False, engine_rounding=False, emulate_upconv=True,
                 encode_only=False):
    """Forward of ``Unet(**cfg)`` holding ``state`` (a state dict) on input ``x``
    ``[N, input_nc, D, H, W]`` fp32.  Returns the output, or ``(output, taps)``
    when ``layers`` (module indices) is non-empty (network.py:475-529); with
    ``encode_only`` the walk stops at slot ``layers[-1]`` and only the taps come back
    (network.py:521-526).  The reference's ReLU / LeakyReLU / SELU are in-place modules
    (network.py:171-204), so a tensor tapped at the slot right before one of them is
    overwritten by it: such a tap holds the POST-activation values."""
    c = dict(DEFAULTS)
    c.update(cfg)
    assert c["dimension"] == 3 and c["pad_type"] == "reflect"
    assert not c["residual_connection"]
    prog, enc_idx, dec_idx = layer_program(
        c["input_nc"], c["output_nc"], c["num_downs"], c["ngf"], c["norm"],
        c["activation"], c["final_act"], c["doubleconv"], c["use_skip_connection"])
    g = lambda k: torch.as_tensor(state[k]).to(torch.float32)
    sdt = engine_storage_dtype(c)
    _rnd = lambda t: t.to(sdt).to(torch.float32)
    feat, skips, taps = x.to(torch.float32), [], []
    tap_slot = {}                                    # module index -> position in `taps`
    eps = c["norm_eps"]
    n_conv = 0
    # the engine runs the conv behind the LAST nearest upsample as low-resolution + skip halves
    split_last = (engine_rounding and emulate_upconv and not training and c["interp"] == "nearest"
                  and c["norm"] != "instance" and c["ngf"] == 16)
    pending_low = None
    for pos, (op, idx, ci, co) in enumerate(prog):
        if op == "conv":
            w = g(f"model.{idx}.weight")
            b = g(f"model.{idx}.bias") if c["norm"] == "instance" else None   # :292
            nxt = prog[pos + 1][0] if pos + 1 < len(prog) else ""
            if engine_rounding:
                # engine folds eval-BN into the packed weights before rounding
                if nxt == "norm" and c["norm"] == "batch" and not training:
                    s = g(f"model.{idx+1}.weight") / torch.sqrt(g(f"model.{idx+1}.running_var") + eps)
                    w = w * s.view(-1, 1, 1, 1, 1)
                    b = g(f"model.{idx+1}.bias") - g(f"model.{idx+1}.running_mean") * s
                if pending_low is not None:
                    cs = feat.shape[1] - pending_low.shape[1]          # skip channels come first
                    zero = torch.zeros(w.shape[0]) if b is None else b
                    part = upconv_lowres(pending_low, w[:, cs:], zero, _rnd)
                    feat = conv3d_reflect(feat[:, :cs], _rnd(w[:, :cs]), None) + part
                    pending_low = None
                    n_conv += 1
                    if c["use_skip_connection"] and idx in enc_idx:
                        skips.append(feat)
                    if idx in layers:
                        tap_slot[idx] = len(taps)
                        taps.append(feat.clone())
                    continue
                if n_conv > 0:                       # stem conv stays fp32
                    w = _rnd(w)
                if nxt == "norm" and c["norm"] == "instance":
                    b = None                         # cancelled by the normalisation; engine drops it
            feat = conv3d_reflect(feat, w, b)
            n_conv += 1
        elif op == "norm":
            if c["norm"] == "batch":                 # network.py:154-155
                if engine_rounding and not training:
                    pass                             # already folded
                else:
                    feat = F.batch_norm(
                        feat, g(f"model.{idx}.running_mean").clone(),
                        g(f"model.{idx}.running_var").clone(),
                        g(f"model.{idx}.weight"), g(f"model.{idx}.bias"),
                        training=training, momentum=0.1, eps=eps)
            elif c["norm"] == "instance":            # network.py:157-158
                if engine_rounding:
                    # statistics from the fp32 accumulators, applied to the stored (rounded) conv output
                    mean = feat.mean(dim=(2, 3, 4), keepdim=True)
                    var = feat.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
                    feat = (_rnd(feat) - mean) * torch.rsqrt(var + eps)
                else:
                    feat = F.instance_norm(feat, eps=eps)
            elif c["norm"] == "instance_affine":
                feat = F.instance_norm(feat, weight=g(f"model.{idx}.weight"),
                                       bias=g(f"model.{idx}.bias"), eps=eps)
        elif op == "act":
            feat = _activate(feat, c["activation"])
            if engine_rounding:
                feat = _rnd(feat)                    # activations live in HBM in the storage type
            if (idx - 1) in tap_slot and c["activation"] in ("relu", "lrelu", "selu"):
                taps[tap_slot[idx - 1]] = feat.clone()   # in-place activation overwrote the tapped tensor
        elif op == "final_act":
            feat = _activate(feat, c["final_act"])
        elif op == "pool":                           # network.py:297,368
            feat = (F.max_pool3d if c["pooling"] == "Max" else F.avg_pool3d)(feat, 2)
            if engine_rounding:
                feat = _rnd(feat)
        elif op == "up":                             # network.py:407
            if split_last and idx == dec_idx[-1]:
                pending_low = feat
            if c["interp"] == "nearest":
                feat = F.interpolate(feat, scale_factor=2, mode="nearest")
            else:
                feat = F.interpolate(feat, scale_factor=2, mode=c["interp"])
                if engine_rounding:
                    feat = _rnd(feat)
        if c["use_skip_connection"]:                 # network.py:543-547
            if idx in dec_idx:
                feat = torch.cat((skips.pop(), feat), dim=1)
            if idx in enc_idx:
                skips.append(feat)
        if idx in layers:
            tap_slot[idx] = len(taps)
            taps.append(feat.clone())
            if encode_only and idx == layers[-1]:
                return taps
    return (feat, taps) if len(layers) else feat


def random_state(cfg, seed, randomize_norm=True):
    """A seeded state dict with the reference's shapes (for cases where no real
    checkpoint exists).  Conv weights ~ U(-k, k) with k = 1/sqrt(fan_in) like
    torch's default init; BatchNorm statistics are made non-trivial so that the
    eval-mode fold is actually exercised."""
    c = dict(DEFAULTS)
    c.update(cfg)
    prog, _, _ = layer_program(
        c["input_nc"], c["output_nc"], c["num_downs"], c["ngf"], c["norm"],
        c["activation"], c["final_act"], c["doubleconv"], c["use_skip_connection"])
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for op, idx, ci, co in prog:
        if op == "conv":
            # Var(w) = 1/fan_in: below the ReLU critical gain (2/fan_in), so a one-ulp
            # rounding difference between two implementations decays instead of being
            # amplified layer after layer; the BatchNorm shifts keep activations O(1)
            k = (3.0 / (ci * 27)) ** 0.5
            sd[f"model.{idx}.weight"] = (torch.rand(co, ci, 3, 3, 3, generator=gen) * 2 - 1) * k
            if c["norm"] == "instance":
                sd[f"model.{idx}.bias"] = (torch.rand(co, generator=gen) * 2 - 1) * 0.1
        elif op == "norm" and c["norm"] == "batch":
            if randomize_norm:
                sd[f"model.{idx}.weight"] = 0.5 + torch.rand(co, generator=gen)
                sd[f"model.{idx}.bias"] = torch.rand(co, generator=gen) - 0.5
                sd[f"model.{idx}.running_mean"] = torch.rand(co, generator=gen) - 0.5
                sd[f"model.{idx}.running_var"] = 0.5 + torch.rand(co, generator=gen)
            else:
                sd[f"model.{idx}.weight"] = torch.ones(co)
                sd[f"model.{idx}.bias"] = torch.zeros(co)
                sd[f"model.{idx}.running_mean"] = torch.zeros(co)
                sd[f"model.{idx}.running_var"] = torch.ones(co)
            sd[f"model.{idx}.num_batches_tracked"] = torch.tensor(0)
        elif op == "norm" and c["norm"] == "instance_affine":
            sd[f"model.{idx}.weight"] = 0.5 + torch.rand(co, generator=gen)
            sd[f"model.{idx}.bias"] = torch.rand(co, generator=gen) - 0.5
    return sd
