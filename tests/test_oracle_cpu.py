"""The CPU oracle (oracle/rtb_oracle.c) against the golden vectors generated from the reference's own
kernels (tests/golden/make_golden.py) and, where /root/reference exists, against the reference itself."""
import os

import numpy as np
import pytest

from common import WALLS_CAMERA, cam12, golden, hex32, mesh_path


def test_camera_basis_and_rays_match_reference(orc):
    g = golden()["camera"]
    assert len(g) >= 5
    for key, rec in g.items():
        W, H = (int(v) for v in key.split("_")[0].split("x"))
        basis = orc.camera_basis(W, H, rec["cam12"])
        assert hex32(basis) == rec["basis"], key
        rays = orc.rays(basis, W, H)
        for i, bits in rec["rays"].items():
            assert hex32(rays[int(i)]) == bits, (key, i)
        assert orc.fnv1a64(rays) == rec["ray_table_hash"], key


def test_default_camera_bit_patterns(orc):
    """SURVEY.md Appendix A.1 known answers for the app's 960x540 camera."""
    b = orc.camera_basis(960, 540, cam12(960, 540))
    assert hex32(b[9:12]) == ["bcae94a4", "bc443e71", "3d6147ae"]          # n_mod
    assert hex32(b[15:16]) == ["383a69dc"] and hex32(b[13:14]) == ["383a69dc"]  # u_mod.x, v_mod.y
    rays = orc.rays(b, 960, 540)
    assert hex32(rays[0]) == ["beb54931", "be4bc7ff", "3f69eebb"]
    assert hex32(rays[259680]) == ["39d3d343", "39d3d589", "3f7ffffc"]


def test_transform_recurrence_matches_reference(orc):
    g = golden()["transform"]
    x = orc.Xform(g["cam_pos"])
    for op, want in zip(g["script"], g["matrices"]):
        x.apply(int(op[0]), *[float(v) for v in op[1:]])
        assert hex32(x.matrix()) == want
    # first R-key step: the camera orbits the origin
    y = orc.Xform([0.0, 0.1, -1.0])
    y.apply(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
    m = y.matrix()
    assert abs(m[3] - 0.198019803) < 1e-8 and m[7] == 0.0 and abs(m[11] + 0.0198019743) < 1e-8


def test_sort_tie_order_is_key_then_descending_index(orc):
    """sort.h:31-54 takes the right run on ties; with all keys equal every split is decided by index."""
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    pts = np.tile(tri, (8, 1))
    nodes = orc.build_tree(pts)
    leaves = nodes[nodes["is_leaf"] == 1]
    assert sorted(leaves["tri"].tolist()) == list(range(8))
    # left subtree of the root gets the first half of the sorted list = highest original indices
    root = nodes[0]
    left_leaf_ids = []
    stack = [int(root["left"])]
    while stack:
        k = stack.pop()
        if nodes[k]["is_leaf"]:
            left_leaf_ids.append(int(nodes[k]["tri"]))
        else:
            stack += [int(nodes[k]["left"]), int(nodes[k]["right"])]
    assert sorted(left_leaf_ids) == [4, 5, 6, 7]


def check_tree_invariants(nodes, n):
    assert len(nodes) == 2 * n - 1
    leaf = nodes["is_leaf"] == 1
    assert leaf.sum() == n and sorted(nodes["tri"][leaf].tolist()) == list(range(n))
    inner = ~leaf
    assert np.array_equal(nodes["right"][inner], nodes["left"][inner] + 1)          # BFS numbering
    L, R = nodes[nodes["left"][inner]], nodes[nodes["right"][inner]]
    P = nodes[inner]
    for lo, hi in (("x0", "x1"), ("y0", "y1"), ("z0", "z1")):                        # exact child bounds
        assert np.array_equal(np.minimum(L[lo], R[lo]), P[lo]) and np.array_equal(np.maximum(L[hi], R[hi]), P[hi])
    ax = P["cut_flag"] % 3
    lmax = np.choose(ax, [L["x1"], L["y1"], L["z1"]])
    rmin = np.choose(ax, [R["x0"], R["y0"], R["z0"]])
    assert np.array_equal(P["s1"], lmax) and np.array_equal(P["s2"], rmin)


def test_tree_invariants_and_golden_hashes(orc, rtb):
    from tests.golden.make_golden import tree_hashes
    pts = rtb.geodesic_mesh(16)
    assert orc.fnv1a64(pts) == golden()["ico16_points_hash"]
    nodes = orc.build_tree(pts)
    check_tree_invariants(nodes, len(pts))
    assert tree_hashes(nodes) == golden()["ico16_tree"]


def test_icosphere_frames_match_reference_golden(orc, rtb):
    g = golden()["ico16_320x180"]["frames"]
    full = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ico16.npz"))
    pts = rtb.geodesic_mesh(16)
    s = orc.Scene(pts, 320, 180, cam12(320, 180))
    n = s.basis[0:3]
    R = (0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
    for k in range(5):
        if k in (1, 2):
            s.transform(10, *R)
        if k == 3:
            for _ in range(100):
                s.transform(32, float(n[0]), float(n[1]), float(n[2]), 0.005)
        if k == 4:
            s.transform(11, 0.0, -0.09950371902099893, 0.0, 0.9950371902099893)
        assert hex32(s.matrix()) == g[k]["matrix"]
        ids, bgra = s.render()
        assert np.array_equal(ids.astype(np.int32), full["ids"][k])
        assert np.array_equal(bgra, full["bgra"][k])
        assert orc.fnv1a64(ids) == g[k]["id_hash"] and orc.fnv1a64(bgra) == g[k]["colour_hash"]
        assert int((ids >= 0).sum()) == g[k]["hits"]


def test_bunny_matches_reference_golden(orc):
    from tests.golden.make_golden import tree_hashes
    path = mesh_path("rabbit_70k.ply")
    if path is None:
        pytest.skip("rabbit_70k.ply not available")
    pts = orc.read_ply(path, 1)
    g = golden()
    assert len(pts) == g["bunny_ply"]["num_tri"] == 69451 and orc.fnv1a64(pts) == g["bunny_ply"]["points_hash"]
    s = orc.Scene(pts, 960, 540, cam12(960, 540))
    assert tree_hashes(s.nodes) == g["bunny_tree"]
    for k in range(2):
        if k:
            s.transform(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
        ids, bgra = s.render()
        f = g["bunny_960x540"]["frames"][k]
        assert (int((ids >= 0).sum()), orc.fnv1a64(ids), orc.fnv1a64(bgra)) == (f["hits"], f["id_hash"], f["colour_hash"])
    # quarter resolution: traversal equals brute force on every pixel (the reference's dormant kernel idea)
    q = orc.Scene(pts, 240, 135, cam12(240, 135), nodes=s.nodes)
    ids, _ = q.render()
    bids, _ = q.render_bruteforce()
    assert np.array_equal(ids >= 0, bids >= 0)


def test_three_walls_matches_reference_golden(orc):
    path = mesh_path("3_walls.ply")
    if path is None:
        pytest.skip("3_walls.ply not available")
    pts = orc.read_ply(path, -1)
    g = golden()["walls_960x540"]
    assert pts.shape == (36, 9) and orc.fnv1a64(pts) == g["points_hash"]
    s = orc.Scene(pts, 960, 540, cam12(960, 540, **WALLS_CAMERA))
    ids, bgra = s.render()
    assert (int((ids >= 0).sum()), orc.fnv1a64(ids), orc.fnv1a64(bgra)) == (g["hits"], g["id_hash"], g["colour_hash"])
    u, c = np.unique(ids[ids >= 0], return_counts=True)
    assert {str(int(a)): int(b) for a, b in zip(u, c)} == g["winners"]


def test_oracle_equals_reference_build_when_present(orc, rtb):
    """Direct comparison with the reference's own kernels (only where oracle/_ref was built)."""
    from oracle import refemu
    if not refemu.available():
        pytest.skip("oracle/_ref/libref_emu.so not built (no /root/reference here)")
    pts = rtb.geodesic_mesh(10, radius=0.09, displacement=0.08, seed=77)
    W, H = 200, 150
    cam = cam12(W, H, pos=(0.2, 0.3, -0.8), look_at=(0.0, 0.1, 0.0), up=(0.0, 1.0, 0.1))
    ref = refemu.RefScene(W, H, cam, points9=pts)
    s = orc.Scene(pts, W, H, cam)
    n = s.basis[0:3]
    ops = [(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)] * 3 + [(32, float(n[0]), float(n[1]), float(n[2]), 0.05)] * 4 + \
          [(31, float(s.basis[6]), float(s.basis[7]), float(s.basis[8]), -0.01), (11, 0.0, -0.09950371902099893, 0.0, 0.9950371902099893)]
    for op in [None] + ops:
        if op:
            ref.transform(*op)
            s.transform(*op)
        assert np.array_equal(ref.matrix().view(np.uint32), s.matrix().view(np.uint32))
        rids, rbgra = ref.render()
        ids, bgra = s.render()
        assert np.array_equal(rids, ids) and np.array_equal(rbgra, bgra)


def test_reference_two_object_sequence_equals_one_object(orc, rtb):
    """WinMain.cpp:152-156 registers TWO objects over the mesh and transforms / renders the first (:188, :212).  In the
    reference's own kernels that sequence draws exactly what a single object draws (the second add_object overwrites
    the camera-side arrays with identical ones, Camera.cpp:156,206): the oracle scene is the one-object scene."""
    from oracle import refemu
    if not refemu.available():
        pytest.skip("oracle/_ref/libref_emu.so not built (no /root/reference here)")
    pts = rtb.geodesic_mesh(8)
    W, H = 160, 90
    two = refemu.RefScene(W, H, cam12(W, H), points9=pts, objects=2)
    s = orc.Scene(pts, W, H, cam12(W, H))
    moved = 0
    for k in range(5):
        if k:
            two.transform(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
            s.transform(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
        ids2, bgra2 = two.render()
        ids, bgra = s.render()
        assert np.array_equal(ids2, ids) and np.array_equal(bgra2, bgra), "frame %d" % k
        if k:
            moved += int((ids != prev).sum())
        prev = ids
    assert moved > 0  # the R key really moves the picture
    refemu.RefScene(W, H, cam12(W, H), points9=pts[:4], objects=1)  # (leaves the driver in its one-object default)


def test_600_step_orbit_recurrence_matches_reference(orc, rtb):
    """BASELINE.json configs[2]: 600 R-key steps.  The quaternion is never renormalised (vector.cpp:38-65), so the matrix
    drifts; the product's host recurrence (rtb::Transform, the matrices every sweep frame is rendered with), the C
    restatement and -- where built -- the reference's own Object::transform (Camera.cu:288-329) must agree bit for bit
    on every one of the 600 steps, and through a block of translations afterwards."""
    from oracle import refemu
    pts = rtb.geodesic_mesh(2)
    W, H = 32, 18
    ref = refemu.RefScene(W, H, cam12(W, H), points9=pts) if refemu.available() else None
    x = orc.Xform([0.0, 0.1, -1.0])
    mesh = rtb.Trixel(pts, require_device=False)
    q = (0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
    from cpp_cuda_raytracer_dev_b200 import lib as L
    import ctypes as C
    # the product's recurrence without a GPU: rtb_transform_sequence_host (host only)
    ops = np.zeros((640, 5), np.float32)
    ops[:600] = (10,) + q
    ops[600:620] = (32, 0.0, 0.0, 1.0, 0.005)
    ops[620:640] = (11, 0.0, -0.09950371902099893, 0.0, 0.9950371902099893)
    got = np.empty((640, 12), np.float32)
    pos = np.array([0.0, 0.1, -1.0], np.float32)
    assert L.rtb_transform_sequence_host(pos.ctypes.data, 640, ops.ctypes.data, got.ctypes.data) == 0
    for k in range(640):
        op = ops[k]
        x.apply(int(op[0]), float(op[1]), float(op[2]), float(op[3]), float(op[4]))
        assert np.array_equal(x.matrix().view(np.uint32), got[k].view(np.uint32)), "step %d (oracle vs product)" % k
        if ref is not None:
            ref.transform(int(op[0]), float(op[1]), float(op[2]), float(op[3]), float(op[4]))
            assert np.array_equal(ref.matrix().view(np.uint32), got[k].view(np.uint32)), "step %d (reference vs product)" % k
    # the drift is real: after 600 steps the rotation part is off orthonormality by several float epsilons (measured 5.7e-7)
    R = got[599].reshape(3, 4)[:, :3].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() > 2.0 ** -22
    mesh.close()


def test_arithmetic_edge_cases_match_reference_golden(orc):
    """Rays with exactly zero direction components (1/0, 0*inf, 0/0 in the slab test), zero-thickness boxes, coincident
    and degenerate triangles, quarter-turn rotations, the camera plane sliding through box planes: the restatement
    against the hashes recorded from the reference's own kernels (tests/golden/make_golden_edge.py)."""
    from common import AXIS_CAMERA, EDGE_FRAMES, GRID_CAMERA, axis_aligned_soup, edge_script, golden_edge, grid_mesh
    g = golden_edge()
    for W, H in EDGE_FRAMES:
        s = orc.Scene(axis_aligned_soup(), W, H, cam12(W, H, **AXIS_CAMERA))
        for k, op in enumerate(edge_script()):
            if op:
                s.transform(*op)
            ids, bgra = s.render()
            want = g["axis_%dx%d" % (W, H)][k]
            assert int((ids >= 0).sum()) == want["hits"], (W, H, k)
            assert orc.fnv1a64(ids) == want["id_hash"] and orc.fnv1a64(bgra) == want["colour_hash"], (W, H, k)
        s.close()
    for W, H in ((33, 33), (128, 72)):
        s = orc.Scene(grid_mesh(), W, H, cam12(W, H, **GRID_CAMERA))
        ids, bgra = s.render()
        want = g["grid_%dx%d" % (W, H)]
        assert orc.fnv1a64(ids) == want["id_hash"] and orc.fnv1a64(bgra) == want["colour_hash"]
        s.close()


def test_scene_extension_default_is_the_reference_pinned_path(orc, rtb):
    """orc_render_scene (lights, shadows, sample_rate, several objects -- defined here because the
    reference leaves them dormant) must leave today's output untouched: one object, the light of Camera.cu:32, no shadows,
    one ray per pixel == orc_render (pinned to the reference), on a moved object and on the all-ties scene."""
    pts = rtb.geodesic_mesh(12)
    W, H = 200, 120
    s = orc.Scene(pts, W, H, cam12(W, H))
    for k in range(3):
        s.transform(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
        a, b = s.render(), orc.render_scene([s])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        c = orc.render_scene([s], sample_rate=1)
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    path = mesh_path("3_walls.ply")
    if path is not None:
        w = orc.Scene(orc.read_ply(path, -1), 320, 180, cam12(320, 180, **WALLS_CAMERA))
        a, b = w.render(), orc.render_scene([w])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and (a[0] >= 0).sum() > 1000
        # two coincident copies of the scene: every hit is a tie between the objects, the first registered one wins
        w2 = orc.Scene(orc.read_ply(path, -1), 320, 180, cam12(320, 180, **WALLS_CAMERA))
        c = orc.render_scene([w, w2])
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])


def test_scene_extension_matches_its_fixtures_and_makes_sense(orc):
    """The extension's definition against the hashes recorded from it (tests/golden/make_golden_scene.py), plus properties
    that do not depend on fixtures: shadows never change hit ids and only darken; a second, farther copy of an object
    changes nothing; a nearer object hides the farther one exactly where it is hit."""
    from common import build_scene_case, golden_scene, scene_cases
    from oracle import standin
    g = golden_scene()
    for name, case in scene_cases().items():
        scenes, kw = build_scene_case(orc, standin.geodesic_mesh, case)
        ids, bgra = orc.render_scene(scenes, **kw)
        want = g[name]
        assert int((ids >= 0).sum()) == want["hits"] and orc.fnv1a64(ids) == want["id_hash"] and orc.fnv1a64(bgra) == want["colour_hash"], name
        if kw["shadows"] and kw["sample_rate"] < 2:
            ids0, bgra0 = orc.render_scene(scenes, lights=kw["lights"], shadows=False)
            assert np.array_equal(ids, ids0), name
            changed = bgra != bgra0
            assert changed.any() and (ids[changed] >= 0).all(), name      # only hit pixels change
        if len(scenes) == 2 and name == "two_objects":
            a, _ = orc.render_scene(scenes[:1]); b, _ = orc.render_scene(scenes[1:])
            both = (a >= 0) & (b >= 0)
            assert both.any()
            merged = np.where(ids >= scenes[0].n, ids - scenes[0].n, ids)
            assert np.array_equal(merged[(a >= 0) & (b < 0)], a[(a >= 0) & (b < 0)]) and np.array_equal(merged[(b >= 0) & (a < 0)], b[(b >= 0) & (a < 0)])
            assert ((ids >= 0) == ((a >= 0) | (b >= 0))).all()
        for s in scenes:
            s.close()
