#!/usr/bin/env python3
"""Fixtures of the scene extension (lights, shadow rays, sample_rate, several objects).

The reference has no behaviour for these (its light loop, shadow test, sample_rate and second object are dormant), so there
is nothing of the reference's to record: "parity unpinned" for this extension.  What is recorded here are the outputs of
the DEFINITION, oracle/rtb_oracle.c (orc_render_scene), so that (a) the oracle itself cannot drift unnoticed
(tests/test_oracle_cpu.py) and (b) the GPU tests have hashes that travel to the box.  The default case of every scene
(one object, light (2,2,2), no shadows, one ray per pixel) is NOT taken from here but from the reference-pinned path.

    python tests/golden/make_golden_scene.py     -> tests/golden/golden_scene.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import orc, standin
from common import cam12, scene_cases, build_scene_case


def main():
    out = {}
    for name, case in scene_cases().items():
        scenes, kw = build_scene_case(orc, standin.geodesic_mesh, case)
        ids, bgra = orc.render_scene(scenes, **kw)
        out[name] = {"hits": int((ids >= 0).sum()), "id_hash": orc.fnv1a64(ids), "colour_hash": orc.fnv1a64(bgra),
                     "black_hit_pixels": int(((bgra == 0) & (ids >= 0)).sum()),
                     "second_object_hits": int((ids >= scenes[0].n).sum()) if len(scenes) > 1 else 0}
        for s in scenes:
            s.close()
    with open(os.path.join(HERE, "golden_scene.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
