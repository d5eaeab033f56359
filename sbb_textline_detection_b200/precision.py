"""Per-layer precision planner (VERDICT r1 task 9).

The parity mode carries every operand as an fp16 (hi, lo) pair and issues hi*hi + hi*lo + lo*hi: three MMA units
per K step, which caps the algorithmic tensor-core fraction at 1/3.  Whether a layer needs the lo plane of its
ACTIVATIONS is a property of the weights: dropping it costs 2^-12 relative rounding on that layer's input, and how
much of that reaches the logits depends on what the rest of the network does with it.  The randomly initialised
synthetic models amplify every layer's rounding beyond the 1e-3 logit tolerance (tools/sim_precision.py); a model
whose logits are dominated by a few well-conditioned paths tolerates hi-only activations in most layers.

``plan_layers`` measures instead of guessing:
  1. logits of sample tiles with the full scheme (the reference the plan is judged against);
  2. for every conv launch alone in hi-only mode, the max |delta logit| it causes;
  3. layers are admitted in order of increasing damage while the MEASURED error of the cumulative plan stays within
     ``budget`` (errors do not add linearly, so every admission is verified by a forward, not predicted).
The result is installed with ``SbbModel.set_precision_plan`` (C ABI: sbb_model_set_precision_plan); hi-only launches
skip the A_lo loads and the A_lo x B_hi product: 2 MMA units per K step.
"""
from __future__ import annotations

import numpy as np


def conv_layers(model):
    return [name for name, _ms, flops in model.layer_times() if flops > 0]


def layer_damage(model, tiles, ref_logits=None):
    """-> (ref_logits, {layer: max |delta logit| when only that layer runs hi-only})"""
    model.set_precision_plan(())
    if ref_logits is None:
        ref_logits = model.predict_tiles(tiles, False, False, True)[2]
    damage = {}
    for name in conv_layers(model):
        model.set_precision_plan((name,))
        z = model.predict_tiles(tiles, False, False, True)[2]
        damage[name] = float(np.abs(z - ref_logits).max())
    model.set_precision_plan(())
    return ref_logits, damage


def plan_layers(model, tiles, budget: float = 4e-4, ref_logits=None, verify_every: int = 1):
    """Greedy plan: the cheapest-to-drop layers first, each admission verified by a forward of the whole plan.
    ``budget`` is the max |delta logit| the plan may add on ``tiles`` relative to the full scheme (the full scheme
    itself sits 3-6e-4 from the fp32 oracle, so the default leaves the 1e-3 tolerance intact).
    Returns (plan tuple, measured error of the plan, per-layer damage dict); the plan is left installed."""
    ref_logits, damage = layer_damage(model, tiles, ref_logits)
    plan, err = [], 0.0
    for name in sorted(damage, key=damage.get):
        if damage[name] == 0.0:
            continue      # no effect at all: the layer has no separate lo operand (conv1 reads packed (hi, lo) pixels)
        if damage[name] > budget:
            break
        trial = plan + [name]
        if len(trial) % verify_every == 0 or damage[name] > 0.25 * budget:
            model.set_precision_plan(trial)
            e = float(np.abs(model.predict_tiles(tiles, False, False, True)[2] - ref_logits).max())
            if e > budget:
                continue
            err = e
        plan = trial
    model.set_precision_plan(plan)
    err = float(np.abs(model.predict_tiles(tiles, False, False, True)[2] - ref_logits).max())
    if err > budget:          # unverified admissions (verify_every > 1) pushed it over: fall back to verifying each
        return plan_layers(model, tiles, budget, ref_logits, verify_every=1)
    return tuple(plan), err, damage


def mma_units(model, plan):
    """Issued MMA units per algorithmic MMA over the whole network: 3 for the full scheme, 2 for hi-only layers
    (FLOP-weighted) -- the ceiling of the algorithmic tensor-core fraction is 1 / this."""
    tot = w = 0.0
    for name, _ms, flops in model.layer_times():
        tot += flops
        w += flops * (2.0 if name in plan else 3.0)
    return w / tot
