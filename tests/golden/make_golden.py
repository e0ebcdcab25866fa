#!/usr/bin/env python3
"""Generate tests/golden/golden.json (+ golden_ico16.npz) from the REFERENCE ITSELF.

Runs only where /root/reference exists: the reference's own kernels, compiled for the host by
oracle/build_ref.py (oracle/_ref/libref_emu.so), are executed on the inputs below and their
outputs are recorded as known-answer vectors.  The C restatement (oracle/rtb_oracle.c) and the
CUDA path are then tested against these files everywhere, including on the GPU box where the
reference does not exist.  Hashes are 64-bit FNV-1a over the raw little-endian buffers
(hit ids as int64, colours as uint32 0x00RRGGBB).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import WALLS_CAMERA, cam12, hex32  # noqa: E402
from oracle import build_ref, orc, refemu  # noqa: E402
import cpp_cuda_raytracer_dev_b200 as rtb  # noqa: E402  (mesh generator + loaders: input preparation only)

REF_DIR = "/root/reference/TEST_Dungeonrun"
R_KEY = (0.0, 0.09950371902099893, 0.0, 0.9950371902099893)


def frame_record(ids, bgra):
    return dict(hits=int((ids >= 0).sum()), id_hash=orc.fnv1a64(ids), colour_hash=orc.fnv1a64(bgra))


def tree_hashes(nodes):
    leaf = nodes["is_leaf"] == 1
    out = {}
    out["left"] = orc.fnv1a64(np.where(leaf, -1, nodes["left"]).astype(np.int64))
    out["right"] = orc.fnv1a64(np.where(leaf, -1, nodes["right"]).astype(np.int64))
    out["tri"] = orc.fnv1a64(np.where(leaf, nodes["tri"], -1).astype(np.int64))
    out["cut_flag"] = orc.fnv1a64(nodes["cut_flag"].astype(np.int32))
    out["bounds"] = orc.fnv1a64(np.stack([nodes[k] for k in ("x0", "x1", "y0", "y1", "z0", "z1")], 1).astype(np.float32))
    out["s1_interior"] = orc.fnv1a64(np.where(leaf, np.float32(0), nodes["s1"]).astype(np.float32))
    out["s2_interior"] = orc.fnv1a64(np.where(leaf, np.float32(0), nodes["s2"]).astype(np.float32))
    out["num_nodes"] = int(len(nodes))
    out["max_depth_leaf_count"] = int(leaf.sum())
    return out


def main():
    build_ref.build()
    G = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libref_emu.so (reference kernels, host build)"}

    # ---- camera bases and primary rays (Camera.cpp:5-67, Camera.cu:89-111) ----------------------
    G["camera"] = {}
    walls_pts = rtb.read_ply(os.path.join(REF_DIR, "3_walls.ply"), -1)
    for (W, H), kw in [((960, 540), {}), ((3840, 2160), {}), ((641, 479), {}), ((320, 180), dict(pos=(0.3, 0.4, -0.9), look_at=(0.0, 0.1, 0.0), up=(0.1, 1.0, 0.0))),
                       ((960, 540), WALLS_CAMERA)]:
        s = refemu.RefScene(W, H, cam12(W, H, **kw), points9=walls_pts[:2])
        rays = s.rays()
        pick = [0, 1, W - 1, W, (H // 2) * W + W // 2, W * H - 1]
        key = "%dx%d%s" % (W, H, "" if not kw else "_" + "_".join("%g" % v for v in kw["pos"]))
        G["camera"][key] = dict(cam12=[float(v) for v in cam12(W, H, **kw)], basis=hex32(s.camera()),
                                 rays={str(i): hex32(rays[i]) for i in pick}, ray_table_hash=orc.fnv1a64(rays))

    # ---- transform recurrence (Camera.cu:254-335) -------------------------------------------------
    s = refemu.RefScene(64, 36, cam12(64, 36), points9=walls_pts[:2])
    n, u = s.camera()[0], s.camera()[2]
    script = []
    for k in range(48):
        if k % 7 == 3:
            script.append((32, float(n[0]), float(n[1]), float(n[2]), 0.005 * (1 + k % 3)))
        elif k % 11 == 5:
            script.append((31, float(u[0]), float(u[1]), float(u[2]), -0.005))
        elif k % 13 == 8:
            script.append((11, 0.0, -0.09950371902099893, 0.0, 0.9950371902099893))
        else:
            script.append((10,) + R_KEY)
    mats = []
    for op in script:
        s.transform(*op)
        mats.append(hex32(s.matrix()))
    G["transform"] = dict(cam_pos=[0.0, 0.1, -1.0], script=[list(op) for op in script], matrices=mats)

    # ---- bunny (BASELINE.json configs[1]) ----------------------------------------------------------
    bunny = os.path.join(REF_DIR, "rabbit_70k.ply")
    s = refemu.RefScene(960, 540, cam12(960, 540), ply_path=bunny, mode=1)
    pts = s.points()
    G["bunny_ply"] = dict(num_tri=int(len(pts)), points_hash=orc.fnv1a64(pts))
    G["bunny_tree"] = tree_hashes(s.nodes())
    frames = []
    for k in range(4):
        if k:
            s.transform(10, *R_KEY)
        ids, bgra = s.render()
        rec = frame_record(ids, bgra)
        rec["matrix"] = hex32(s.matrix())
        frames.append(rec)
    G["bunny_960x540"] = dict(frames=frames)
    s = refemu.RefScene(960, 540, cam12(960, 540), ply_path=bunny, mode=1)
    n = s.camera()[0]
    for _ in range(150):
        s.transform(32, float(n[0]), float(n[1]), float(n[2]), 0.005)
    ids, bgra = s.render()
    G["bunny_960x540_closeup"] = dict(frame_record(ids, bgra), matrix=hex32(s.matrix()))
    s = refemu.RefScene(480, 270, cam12(480, 270), ply_path=bunny, mode=1)
    ids, bgra = s.render()
    G["bunny_480x270"] = frame_record(ids, bgra)

    # ---- 3_walls (BASELINE.json configs[0]) ---------------------------------------------------------
    s = refemu.RefScene(960, 540, cam12(960, 540, **WALLS_CAMERA), points9=walls_pts)
    ids, bgra = s.render()
    uniq, cnt = np.unique(ids[ids >= 0], return_counts=True)
    G["walls_960x540"] = dict(frame_record(ids, bgra), winners={str(int(a)): int(b) for a, b in zip(uniq, cnt)},
                              points_hash=orc.fnv1a64(walls_pts), tree=tree_hashes(s.nodes()))

    # ---- procedural icosphere (travels without any asset) -----------------------------------------
    ico = rtb.geodesic_mesh(16)
    G["ico16_points_hash"] = orc.fnv1a64(ico)
    s = refemu.RefScene(320, 180, cam12(320, 180), points9=ico)
    G["ico16_tree"] = tree_hashes(s.nodes())
    n = s.camera()[0]
    frames, keep_ids, keep_col = [], [], []
    for k in range(5):
        if k in (1, 2):
            s.transform(10, *R_KEY)
        if k == 3:
            for _ in range(100):
                s.transform(32, float(n[0]), float(n[1]), float(n[2]), 0.005)
        if k == 4:
            s.transform(11, 0.0, -0.09950371902099893, 0.0, 0.9950371902099893)
        ids, bgra = s.render()
        rec = frame_record(ids, bgra)
        rec["matrix"] = hex32(s.matrix())
        frames.append(rec)
        keep_ids.append(ids.astype(np.int32))
        keep_col.append(bgra)
    G["ico16_320x180"] = dict(frames=frames)
    np.savez_compressed(os.path.join(HERE, "golden_ico16.npz"), ids=np.stack(keep_ids), bgra=np.stack(keep_col))

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote golden.json and golden_ico16.npz")


if __name__ == "__main__":
    main()
