"""Flat layer plan of the anatomix 3-D U-Net.

The reference builds its network as one flat ``nn.Sequential`` whose *indices*
are part of the public contract (state-dict keys ``model.<idx>.*``, the skip
bookkeeping lists ``encoder_idx`` / ``decoder_idx`` and the feature-tap indices
callers pass to ``forward(layers=...)``), see reference
``anatomix/model/network.py:309-465``.  This module derives that index layout
from the constructor arguments as plain data (no torch), so that the nn.Module
mirror (`anatomix_b200.unet`), the B200 engine planner (`anatomix_b200.engine`)
and the CPU oracle all agree on it.

Slot kinds: ``conv``, ``norm``, ``act``, ``pool``, ``up``, ``final_act``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass(frozen=True)
class Slot:
    kind: str                 # conv | norm | act | pool | up | final_act
    index: int                # position in the flat Sequential
    cin: int = 0              # conv only
    cout: int = 0             # conv / norm feature count
    level: int = 0            # resolution level (0 = full res, k = 1/2^k)
    role: str = ""            # stem | enc | bott | dec | final


@dataclass
class UnetPlan:
    """Everything index-related the reference constructor decides."""
    slots: List[Slot] = field(default_factory=list)
    encoder_idx: List[int] = field(default_factory=list)   # network.py:367
    decoder_idx: List[int] = field(default_factory=list)   # network.py:406
    res_source: List[int] = field(default_factory=list)    # network.py:320 ...
    res_dest: List[int] = field(default_factory=list)      # network.py:326 ...

    @property
    def convs(self) -> List[Slot]:
        return [s for s in self.slots if s.kind == "conv"]


def make_plan(input_nc: int, output_nc: int, num_downs: int, ngf: int,
              has_norm: bool = True, has_act: bool = True,
              has_final_act: bool = False, doubleconv: bool = True,
              use_skip_connection: bool = True) -> UnetPlan:
    """Index layout of ``Unet(...)`` (reference network.py:309-465).

    stem conv; per encoder level one or two convs then a pool (level 0 keeps the
    width, deeper levels double it); a one/two-conv bottleneck that doubles the
    width again; per decoder level an upsample, a conv that eats
    ``[skip | upsampled]`` and halves the width, and an optional second conv;
    a last conv to ``output_nc`` with no norm.
    """
    plan = UnetPlan()

    def push(kind, **kw):
        s = Slot(kind=kind, index=len(plan.slots), **kw)
        plan.slots.append(s)
        return s.index

    def conv_block(cin, cout, level, role):
        plan.res_source.append(push("conv", cin=cin, cout=cout, level=level, role=role))
        if has_norm:
            push("norm", cout=cout, level=level, role=role)
        if has_act:
            push("act", cout=cout, level=level, role=role)
        # reference records len(model)-1 whatever the last slot was (network.py:326)
        plan.res_dest.append(len(plan.slots) - 1)

    conv_block(input_nc, ngf, 0, "stem")
    width = ngf
    for lvl in range(num_downs):
        grown = width if lvl == 0 else 2 * width
        conv_block(width, grown, lvl, "enc")
        if doubleconv:
            conv_block(grown, grown, lvl, "enc")
        plan.encoder_idx.append(len(plan.slots) - 1)
        push("pool", cout=grown, level=lvl, role="enc")
        width = grown

    conv_block(width, 2 * width, num_downs, "bott")
    if doubleconv:
        conv_block(2 * width, 2 * width, num_downs, "bott")

    mult = 2 ** num_downs
    for j in range(num_downs):
        lvl = num_downs - 1 - j
        plan.decoder_idx.append(push("up", cout=ngf * mult, level=lvl, role="dec"))
        fan_in = mult + mult // 2 if use_skip_connection else mult
        half = ngf * (mult // 2)
        conv_block(ngf * fan_in, half, lvl, "dec")
        if doubleconv:
            conv_block(half, half, lvl, "dec")
        mult //= 2

    push("conv", cin=ngf * mult, cout=output_nc, level=0, role="final")
    if has_final_act:
        push("final_act", cout=output_nc, level=0, role="final")
    return plan
