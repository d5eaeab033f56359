"""Per-layer precision plan of a model (sbb_textline_detection_b200/precision.py) on sample tiles of a synthetic page:
    python tools/precision_plan.py [--weights random|semantic] [--kind textline] [--tile 448] [--budget 4e-4]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sbb_textline_detection_b200 import precision, semantic, synth  # noqa: E402
from sbb_textline_detection_b200.detector import synthetic_weights  # noqa: E402
from sbb_textline_detection_b200.model import SbbModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--weights", default="random", choices=["random", "semantic"])
ap.add_argument("--kind", default="textline")
ap.add_argument("--tile", type=int, default=448)
ap.add_argument("--budget", type=float, default=4e-4)
a = ap.parse_args()
w, nc = synthetic_weights(a.kind) if a.weights == "random" else semantic.semantic_weights(a.kind)
T = a.tile
page = synth.document_page(2800, 2000, seed=0)
tiles = np.stack([page[360:360 + T, 360:360 + T], page[1200:1200 + T, 900:900 + T], synth.uniform_page(T, T, 0)])
tiles = tiles.astype(np.float32) / np.float32(255)
m = SbbModel(w, T, T, nc, max_batch=4)
m.predict_tiles(tiles, False, False, True)          # fills layer_times' FLOP column
plan, err, damage = precision.plan_layers(m, tiles, budget=a.budget)
print(json.dumps({"weights": a.weights, "kind": a.kind, "tile": T, "budget": a.budget, "plan_layers": len(plan),
                  "of": len(damage), "plan_error": err, "mma_units_per_alg_mma": precision.mma_units(m, plan),
                  "damage_min": min(damage.values()), "damage_median": float(np.median(list(damage.values()))),
                  "damage_max": max(damage.values())}))
for name in sorted(damage, key=damage.get):
    print(f"  {name:18s} {damage[name]:.3e} {'hi-only' if name in plan else ''}")
m.close()
