"""Mints the golden vectors in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``anatomix.model.network.Unet`` from ``/root/reference`` (with this
repo's ``anatomix`` shim kept off ``sys.path``), loads the released 6M checkpoint
``model-weights/anatomix.pth`` and records reference outputs for the cases of
SURVEY.md appendix E.  The reference has no tests of its own, so these files are
the pins for ``oracle/`` (tests/test_oracle.py) and, through the oracle, for the
CUDA engine.  Large tensors are stored as strided samples plus moments.

Also exports the checkpoint as ``anatomix_6m_state.npz`` (plain fp32 arrays) so
that GPU boxes without /root/reference can run parity on the real weights.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CFG_6M = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
CFG_94M = dict(dimension=3, input_nc=1, output_nc=32, num_downs=5, ngf=32,
               norm="instance", pooling="Avg", interp="trilinear", norm_eps=1e-2)


def rand_input(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def moments(t):
    t = t.double()
    return np.array([t.mean().item(), t.std().item(), t.min().item(), t.max().item(),
                     t.abs().sum().item()], dtype=np.float64)


def sub(t, step):
    return t[:, :, ::step, ::step, ::step].contiguous().numpy()


def structured_inputs(S=32):
    z = torch.zeros(1, 1, S, S, S)
    out = {"zeros": z.clone(), "ones": torch.ones_like(z)}
    for name, pos in (("corner", (0, 0, 0)), ("face", (0, S // 2, S // 2)),
                      ("centre", (S // 2, S // 2, S // 2)), ("far_corner", (S - 1, S - 1, S - 1))):
        t = z.clone(); t[0, 0, pos[0], pos[1], pos[2]] = 1.0
        out["impulse_" + name] = t
    r = torch.arange(S, dtype=torch.float32) / S
    out["ramp"] = (r[:, None, None] * 0.5 + r[None, :, None] * 0.3 + r[None, None, :] * 0.2)[None, None].contiguous()
    return out


def write_manifest():
    with open(os.path.join(HERE, "MANIFEST.txt"), "w") as f:
        f.write("minted by tests/golden/make_golden.py from /root/reference (torch %s)\n" % torch.__version__)
        for name in sorted(os.listdir(HERE)):
            if name.endswith(".npz"):
                h = hashlib.sha256(open(os.path.join(HERE, name), "rb").read()).hexdigest()[:16]
                f.write(f"{name} {os.path.getsize(os.path.join(HERE, name))} {h}\n")
    print(open(os.path.join(HERE, "MANIFEST.txt")).read())


def mint_g7(Unet, sd):
    """G7: the feature-tap branch (network.py:475-529) at slots the engine serves -- norm slots (which hold
    POST-activation values: the in-place ReLU overwrites the tapped tensor), activation, pooling and
    Upsample (concat) slots, the last conv -- plus an `encode_only` call."""
    m6 = Unet(**CFG_6M); m6.load_state_dict(sd, strict=True); m6.eval()
    ids = [1, 2, 9, 16, 30, 36, 37, 44, 51, 58, 60, 61, 64, 65]
    enc = [8, 22, 15]                      # stops at slot layers[-1] = 15: taps of 8 and 15 only
    with torch.no_grad():
        x = rand_input((1, 1, 32, 32, 32), 3)
        y, taps = m6(x, layers=ids)
        g = {"tap_ids": np.array(ids), "out_s2": sub(y, 2), "enc_ids": np.array(enc)}
        for i, t in zip(ids, taps):
            g[f"tap{i}"] = sub(t, 2) if t.shape[-1] > 4 else t.numpy()
        only = m6(x, layers=enc, encode_only=True)
        g["enc_count"] = np.array(len(only))
        for k, t in enumerate(only):
            g[f"enc{k}"] = sub(t, 2)
    np.savez_compressed(os.path.join(HERE, "g7_6m_taps.npz"), **g)


def mint_g8(Unet):
    """G8: the 94M `anatomix-dev` config at the BASELINE volume size (1 x 128^3), seeded default init (seed 0, as G4);
    G9: the same network at 64^3 with every conv weight scaled x30 -- the raw pre-norm conv outputs, which the engine
    stores in fp16 before normalising them, are 30x larger there (InstanceNorm undoes the scale up to its eps)."""
    torch.manual_seed(0)
    m94 = Unet(**CFG_94M).eval()
    with torch.no_grad():
        y = m94(rand_input((1, 1, 128, 128, 128), 0))
        np.savez_compressed(os.path.join(HERE, "g8_94m_128.npz"), out_s8=sub(y, 8), mom=moments(y),
                            probe=y[0, :4, 64, 64, 64].numpy())
        for k, v in m94.state_dict().items():
            if k.endswith(".weight"):
                v.mul_(30.0)
        y = m94(rand_input((1, 1, 64, 64, 64), 0))
        np.savez_compressed(os.path.join(HERE, "g9_94m_64_w30.npz"), out_s4=sub(y, 4), mom=moments(y))


def mint_g10(Unet, sd):
    """G10: PRE-norm conv taps (network.py:504-515: `layers` holding conv slots, as the pretraining code does): the
    stem, encoder / bottleneck / decoder convs at every level incl. the level-0 decoder conv the engine may run as two
    launches, for the 6M model; plus the 94M config's stem and a deep conv (InstanceNorm: bias-carrying convs)."""
    m6 = Unet(**CFG_6M); m6.load_state_dict(sd, strict=True); m6.eval()
    ids = [0, 3, 6, 10, 27, 34, 38, 52, 59, 62]
    with torch.no_grad():
        x = rand_input((1, 1, 32, 32, 32), 5)
        y, taps = m6(x, layers=ids)
        g = {"tap_ids": np.array(ids), "out_s2": sub(y, 2)}
        for i, t in zip(ids, taps):
            g[f"tap{i}"] = sub(t, 2) if t.shape[-1] > 4 else t.numpy()
    torch.manual_seed(0)
    m94 = Unet(**CFG_94M).eval()
    ids94 = [0, 3, 38, 76]
    with torch.no_grad():
        x = rand_input((1, 1, 64, 64, 64), 0)
        y, taps = m94(x, layers=ids94)
        g["tap_ids_94m"] = np.array(ids94)
        for i, t in zip(ids94, taps):
            g[f"m94_tap{i}"] = sub(t, 4) if t.shape[-1] > 4 else t.numpy()
    np.savez_compressed(os.path.join(HERE, "g10_prenorm_taps.npz"), **g)


def main():
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(HERE, "..", ".."))]
    sys.path.insert(0, REF)
    from anatomix.model.network import Unet
    import anatomix.model.network as net
    assert net.__file__.startswith(REF), net.__file__
    torch.set_num_threads(os.cpu_count())
    if sys.argv[1:] == ["g10"]:            # add G10 without re-minting the others
        mint_g10(Unet, torch.load(os.path.join(REF, "model-weights", "anatomix.pth"), map_location="cpu"))
        write_manifest()
        return
    if sys.argv[1:] == ["g8"]:             # add G8 / G9 without re-minting the others
        mint_g8(Unet)
        write_manifest()
        return
    if sys.argv[1:] == ["g7"]:             # add G7 without re-minting the others
        mint_g7(Unet, torch.load(os.path.join(REF, "model-weights", "anatomix.pth"), map_location="cpu"))
        write_manifest()
        return

    sd = torch.load(os.path.join(REF, "model-weights", "anatomix.pth"), map_location="cpu")
    np.savez_compressed(os.path.join(HERE, "anatomix_6m_state.npz"),
                        **{k: v.numpy() for k, v in sd.items()})

    m6 = Unet(**CFG_6M); m6.load_state_dict(sd, strict=True); m6.eval()
    taps_6m = [0, 8, 9, 36, 37, 58, 62]
    with torch.no_grad():
        # G1: smallest legal size, full output + sampled taps
        x = rand_input((1, 1, 32, 32, 32), 0)
        y, taps = m6(x, layers=taps_6m)
        g = {"out": y.numpy(), "mom": moments(y), "tap_ids": np.array(taps_6m)}
        for i, t in zip(taps_6m, taps):
            g[f"tap{i}"] = sub(t, 2) if t.shape[-1] > 4 else t.numpy()
            g[f"tap{i}_mom"] = moments(t)
        np.savez_compressed(os.path.join(HERE, "g1_6m_32.npz"), **g)

        # G2: batch > 1, non-cubic
        x = rand_input((2, 1, 32, 48, 32), 1)
        y = m6(x)
        np.savez_compressed(os.path.join(HERE, "g2_6m_2x32x48x32.npz"), out_s2=sub(y, 2), mom=moments(y))

        # G3: headline shape, samples + moments
        x = rand_input((1, 1, 128, 128, 128), 0)
        y = m6(x)
        np.savez_compressed(os.path.join(HERE, "g3_6m_128.npz"), out_s8=sub(y, 8), mom=moments(y),
                            probe=y[0, :4, 64, 64, 64].numpy())

        # G6: structured inputs
        g = {}
        for name, t in structured_inputs().items():
            y = m6(t)
            g[name] = sub(y, 2); g[name + "_mom"] = moments(y)
        np.savez_compressed(os.path.join(HERE, "g6_6m_structured.npz"), **g)

    # G5: train mode (batch statistics) must stay on the torch path
    m6t = Unet(**CFG_6M); m6t.load_state_dict(sd, strict=True); m6t.train()
    with torch.no_grad():
        y = m6t(rand_input((1, 1, 32, 32, 32), 0))
    np.savez_compressed(os.path.join(HERE, "g5_6m_train.npz"), out_s2=sub(y, 2), mom=moments(y))

    # G4: 94M config with seeded default init (the released weights are Hub-only)
    torch.manual_seed(0)
    m94 = Unet(**CFG_94M).eval()
    fp = {k: np.array([v.double().sum().item(), v.double().abs().sum().item()])
          for k, v in m94.state_dict().items()}
    taps_94 = [0, 8, 9, 43, 44, 79]
    with torch.no_grad():
        x = rand_input((1, 1, 64, 64, 64), 0)
        y, taps = m94(x, layers=taps_94)
    g = {"out_s2": sub(y, 2), "mom": moments(y), "tap_ids": np.array(taps_94),
         "param_names": np.array(list(fp)), "param_sums": np.stack(list(fp.values()))}
    for i, t in zip(taps_94, taps):
        g[f"tap{i}"] = sub(t, 4) if t.shape[-1] > 4 else t.numpy()
        g[f"tap{i}_mom"] = moments(t)
    np.savez_compressed(os.path.join(HERE, "g4_94m_64.npz"), **g)

    mint_g7(Unet, sd)
    mint_g8(Unet)
    mint_g10(Unet, sd)
    write_manifest()


if __name__ == "__main__":
    main()
