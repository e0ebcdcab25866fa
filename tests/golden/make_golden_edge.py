#!/usr/bin/env python3
"""Generate tests/golden/golden_edge.json from the REFERENCE ITSELF (oracle/_ref/libref_emu.so, the reference's own
kernels compiled for the host): the arithmetic edge-case scenes of tests/common.py -- rays with exactly zero direction
components, zero-thickness boxes, coincident and degenerate triangles, quarter-turn rotations, a camera plane sliding
through box planes.  Runs only where /root/reference exists; the file travels to the GPU box.

    python tests/golden/make_golden_edge.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import AXIS_CAMERA, EDGE_FRAMES, GRID_CAMERA, axis_aligned_soup, cam12, edge_script, grid_mesh  # noqa: E402
from oracle import build_ref, orc, refemu  # noqa: E402


def record(ids, bgra):
    return dict(hits=int((ids >= 0).sum()), id_hash=orc.fnv1a64(ids), colour_hash=orc.fnv1a64(bgra))


def main():
    build_ref.build()
    G = {"generator": "tests/golden/make_golden_edge.py", "source": "oracle/_ref/libref_emu.so (reference kernels, host build)"}
    for W, H in EDGE_FRAMES:
        ref = refemu.RefScene(W, H, cam12(W, H, **AXIS_CAMERA), points9=axis_aligned_soup())
        frames = []
        for op in edge_script():
            if op:
                ref.transform(*op)
            frames.append(record(*ref.render()))
        G["axis_%dx%d" % (W, H)] = frames
    for W, H in ((33, 33), (128, 72)):
        ref = refemu.RefScene(W, H, cam12(W, H, **GRID_CAMERA), points9=grid_mesh())
        G["grid_%dx%d" % (W, H)] = record(*ref.render())
    with open(os.path.join(HERE, "golden_edge.json"), "w") as f:
        json.dump(G, f, indent=1)
    print("wrote golden_edge.json:", {k: (len(v) if isinstance(v, list) else 1) for k, v in G.items() if k not in ("generator", "source")})


if __name__ == "__main__":
    main()
