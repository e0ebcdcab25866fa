"""Host side of the product (mesh load, tree build, camera, transform; no GPU needed) against the
oracle, and the C-ABI surface of librtb.so."""
import ctypes
import os
import re

import numpy as np
import pytest

from common import ROOT, WALLS_CAMERA, cam12, cam_kwargs, golden, hex32, mesh_path


def test_library_exports_every_declared_symbol(rtb):
    hdr = open(os.path.join(ROOT, "include", "rtb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(rtb.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(rtb.EXPORTS)
    assert rtb.lib.rtb_version().startswith(b"rtb")


def test_status_codes_not_exceptions(rtb):
    """Reference convention: status code + message, never throw."""
    p, n = ctypes.c_void_p(), ctypes.c_uint32()
    rc = rtb.lib.rtb_read_ply(b"/nonexistent/file.ply", 0, ctypes.byref(p), ctypes.byref(n))
    assert rc == 2 and b"cannot read" in rtb.lib.rtb_last_error()
    assert rtb.lib.rtb_mesh_build_tree(None) == 1
    with pytest.raises(rtb.RtbError):
        rtb.read_ply("/nonexistent/file.ply", 0)


def test_ply_loader_matches_oracle(rtb, orc, tmp_path):
    for name, mode in (("rabbit_70k.ply", 1), ("3_walls.ply", -1)):
        path = mesh_path(name)
        if path is None:
            continue
        assert np.array_equal(rtb.read_ply(path, mode).view(np.uint32), orc.read_ply(path, mode).view(np.uint32))
    # the four column modes and both face kinds on a synthetic file
    verts = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1.25]], np.float32)
    for mode, extra in ((0, 0), (1, 2), (2, 3)):
        f = tmp_path / ("m%d.ply" % mode)
        with open(f, "w") as fh:
            fh.write("ply\nformat ascii 1.0\nelement vertex 5\nelement face 3\nend_header\n")
            for v in verts:
                fh.write(" ".join(["%.9g" % c for c in v] + ["0.5"] * extra) + "\n")
            fh.write("4 0 1 2 3\n3 0 1 4\n3 2 3 4\n")
        got = rtb.read_ply(str(f), mode)
        assert np.array_equal(got.view(np.uint32), orc.read_ply(str(f), mode).view(np.uint32))
        assert got.shape == (4, 9)
        assert got[0].tolist() == [0, 0, 0, 1, 0, 0, 1, 1, 0] and got[1].tolist() == [0, 0, 0, 1, 1, 0, 0, 1, 0]   # quad -> (A,B,C),(A,C,D)
        assert got[2].tolist() == [0.5, 0.5, 1.25, 0, 0, 0, 1, 0, 0]                                               # "3 a b c" stored (c,a,b)
    # write_ply round trip
    pts = rtb.geodesic_mesh(3)
    out = tmp_path / "rt.ply"
    rtb.write_ply(str(out), pts)
    assert np.array_equal(rtb.read_ply(str(out), 0).view(np.uint32), pts.view(np.uint32))
    assert np.array_equal(orc.read_ply(str(out), 0).view(np.uint32), pts.view(np.uint32))


def _write_ply(path, fmt, verts, faces, vertex_layout, count_type, index_type, extra_elements=True):
    """A PLY file in any of the three formats with a chosen vertex property layout, for the conforming reader."""
    np_of = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8"}
    end = {"ascii": "=", "binary_little_endian": "<", "binary_big_endian": ">"}[fmt]
    hdr = ["ply", "format %s 1.0" % fmt, "comment made by tests/test_host_cpu.py"]
    if extra_elements:
        hdr += ["element material 2", "property uchar red", "property list uchar float coeffs"]
    hdr += ["element vertex %d" % len(verts)] + ["property %s %s" % (t, n) for t, n in vertex_layout]
    hdr += ["element face %d" % len(faces), "property uchar flags", "property list %s %s vertex_indices" % (count_type, index_type)]
    if extra_elements:
        hdr += ["element edge 1", "property int vertex1", "property int vertex2"]
    hdr += ["end_header"]
    rng = np.random.default_rng(3)
    with open(path, "wb") as fh:
        fh.write(("\n".join(hdr) + "\n").encode())
        def put(values_types):
            if fmt == "ascii":
                fh.write((" ".join(repr(float(v)) if t in ("float", "double") else str(int(v)) for v, t in values_types) + "\n").encode())
            else:
                for v, t in values_types:
                    fh.write(np.array(v, dtype=end + np_of[t]).tobytes())
        if extra_elements:
            put([(7, "uchar"), (2, "uchar"), (0.5, "float"), (0.25, "float")])
            put([(9, "uchar"), (0, "uchar")])
        for v in verts:
            rec = []
            for t, n in vertex_layout:
                rec.append((v["xyz".index(n)] if n in "xyz" else rng.integers(0, 100), t))
            put(rec)
        for f in faces:
            put([(1, "uchar"), (len(f), count_type)] + [(i, index_type) for i in f])
        if extra_elements:
            put([(0, "int"), (1, "int")])


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("layout", ["xyz", "normals_first", "double_xyz"])
def test_conforming_ply_reader(rtb, tmp_path, fmt, layout):
    """Mode -1: header-driven reader -- three formats, any property order and types, list
    properties, extra elements; triangle order rules of read_ply.cpp:92-148 kept (3-gon -> (c,a,b), 4-gon -> fan)."""
    verts = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1.25], [0.25, -1.5, 3.0]], np.float32)
    faces = [[0, 1, 2, 3], [0, 1, 4], [2, 3, 4], [0, 1, 2, 3, 5]]
    vl = {"xyz": [("float", "x"), ("float", "y"), ("float", "z")],
          "normals_first": [("float", "nx"), ("uchar", "red"), ("float", "z"), ("float", "x"), ("short", "s"), ("float", "y"), ("double", "q")],
          "double_xyz": [("double", "x"), ("double", "y"), ("double", "z"), ("float", "confidence")]}[layout]
    ct, it = ("uchar", "int") if layout == "xyz" else ("ushort", "uint") if layout == "normals_first" else ("int", "ushort")
    f = tmp_path / "c.ply"
    _write_ply(str(f), fmt, verts, faces, vl, ct, it)
    got = rtb.read_ply(str(f), -1)
    V = verts
    want = np.array([np.concatenate([V[0], V[1], V[2]]), np.concatenate([V[0], V[2], V[3]]),     # quad -> (A,B,C),(A,C,D)
                     np.concatenate([V[4], V[0], V[1]]), np.concatenate([V[4], V[2], V[3]]),     # (a,b,c) stored (c,a,b)
                     np.concatenate([V[0], V[1], V[2]]), np.concatenate([V[0], V[2], V[3]]), np.concatenate([V[0], V[3], V[5]])], np.float32)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_conforming_reader_equals_reference_reader_on_text_files(rtb, tmp_path):
    """Where both apply (ASCII, x y z first), mode -1 must give what the reference's scanner gives (modes 0/1)."""
    path = mesh_path("rabbit_70k.ply")
    if path is not None:
        # the reference's own text files declare no properties at all (header: counts only), which is why its
        # scanner needs a column mode; a header-driven reader has to refuse them rather than guess
        with pytest.raises(rtb.RtbError):
            rtb.read_ply(path, -1)
    pts = rtb.geodesic_mesh(5)
    out = tmp_path / "rt.ply"
    rtb.write_ply(str(out), pts)
    assert np.array_equal(rtb.read_ply(str(out), -1).view(np.uint32), pts.view(np.uint32))


def test_conforming_reader_rejects_bad_files(rtb, tmp_path):
    f = tmp_path / "bad.ply"
    for body in ("plx\n", "ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n",
                 "ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n\x00\x00",
                 "ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 7\n"):
        f.write_bytes(body.encode("latin1"))
        with pytest.raises(rtb.RtbError):
            rtb.read_ply(str(f), -1)


def tree_equal(t, on):
    leaf = on["is_leaf"] == 1
    ok = np.array_equal(t["left"], np.where(leaf, -1, on["left"]).astype(np.int32))
    ok &= np.array_equal(t["right"], np.where(leaf, -1, on["right"]).astype(np.int32))
    ok &= np.array_equal(t["tri"], np.where(leaf, on["tri"], -1).astype(np.int32))
    ok &= np.array_equal(t["cut_flag"], on["cut_flag"])
    ob = np.stack([on[k] for k in ("x0", "x1", "y0", "y1", "z0", "z1")], 1)
    ok &= np.array_equal(t["bounds"].view(np.uint32), ob.view(np.uint32))
    ok &= np.array_equal(t["s1"][~leaf].view(np.uint32), on["s1"][~leaf].view(np.uint32))
    ok &= np.array_equal(t["s2"][~leaf].view(np.uint32), on["s2"][~leaf].view(np.uint32))
    return bool(ok)


@pytest.mark.parametrize("case", ["ico16", "ico40", "one", "two", "seven", "ties", "bunny", "walls"])
def test_tree_build_is_bit_identical_to_oracle(rtb, orc, case):
    if case == "bunny":
        path = mesh_path("rabbit_70k.ply")
        if path is None:
            pytest.skip("rabbit_70k.ply not available")
        pts = rtb.read_ply(path, 1)
    elif case == "walls":
        path = mesh_path("3_walls.ply")
        if path is None:
            pytest.skip("3_walls.ply not available")
        pts = rtb.read_ply(path, -1)
    elif case == "ties":   # many equal keys: the (key, descending index) order decides every split
        base = rtb.geodesic_mesh(2)
        pts = np.concatenate([base, base, base[::-1]])
    else:
        pts = {"ico16": rtb.geodesic_mesh(16), "ico40": rtb.geodesic_mesh(40), "one": rtb.geodesic_mesh(2)[:1],
               "two": rtb.geodesic_mesh(2)[:2], "seven": rtb.geodesic_mesh(2)[:7]}[case]
    m = rtb.Trixel(pts, require_device=False)
    m.create_kd()
    assert tree_equal(m.tree(), orc.build_tree(pts))
    assert m.num_voxels == 2 * len(pts) - 1
    m.close()


def test_camera_basis_matches_oracle_and_golden(rtb, orc):
    for key, rec in golden()["camera"].items():
        W, H = (int(v) for v in key.split("_")[0].split("x"))
        c = rec["cam12"]
        cam = rtb.Camera(W, H, c[0], c[1], c[2], c[3:6], c[6:9], c[9:12], require_device=False)
        assert hex32(cam.basis()) == rec["basis"], key
        cam.close()


def test_geodesic_mesh_is_deterministic(rtb):
    a, b = rtb.geodesic_mesh(12), rtb.geodesic_mesh(12)
    assert a.shape == (20 * 144, 9) and np.array_equal(a, b)
    assert not np.array_equal(a, rtb.geodesic_mesh(12, seed=5))
    r = np.linalg.norm(a.reshape(-1, 3) - np.array([0.0, 0.1, 0.0]), axis=1)
    assert 0.07 < r.min() and r.max() < 0.09


def test_render_without_gpu_fails_loudly(rtb):
    """No CPU fallback: on a box without a usable device the render path must report an error."""
    if rtb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    pts = rtb.geodesic_mesh(2)
    with pytest.raises(rtb.RtbError):
        rtb.Trixel(pts)                      # device upload fails -> status code -> exception in the binding
    m = rtb.Trixel(pts, require_device=False)
    m.create_kd()
    cam = rtb.Camera(64, 48, **cam_kwargs(64, 48), require_device=False)
    obj = rtb.Object(m)
    with pytest.raises(rtb.RtbError):
        cam.add_object(obj)
    with pytest.raises(rtb.RtbError):
        obj.render(cam)


def test_frame_files(rtb, tmp_path):
    """rtb_write_frame: the headless stand-in of the reference's window blit -- PPM and PNG, bottom-up rows flipped."""
    import struct
    import zlib
    W, H = 5, 3
    frame = (np.arange(W * H, dtype=np.uint32) * 0x010203 + 0x00f08200) & 0x00ffffff
    rgb = np.stack([(frame >> 16) & 0xff, (frame >> 8) & 0xff, frame & 0xff], 1).astype(np.uint8).reshape(H, W, 3)[::-1]
    ppm = tmp_path / "f.ppm"
    rtb.write_frame(str(ppm), frame, W, H)
    data = ppm.read_bytes()
    assert data.startswith(b"P6\n5 3\n255\n") and data[len(b"P6\n5 3\n255\n"):] == rgb.tobytes()
    png = tmp_path / "f.png"
    rtb.write_frame(str(png), frame, W, H)
    data = png.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    at, idat, seen = 8, b"", []
    while at < len(data):
        n, kind = struct.unpack(">I4s", data[at:at + 8])
        body = data[at + 8:at + 8 + n]
        assert struct.unpack(">I", data[at + 8 + n:at + 12 + n])[0] == zlib.crc32(kind + body)
        seen.append(kind)
        if kind == b"IHDR":
            assert struct.unpack(">IIBBBBB", body) == (W, H, 8, 2, 0, 0, 0)
        if kind == b"IDAT":
            idat += body
        at += 12 + n
    assert seen[0] == b"IHDR" and seen[-1] == b"IEND"
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, 1 + 3 * W)
    assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(H, W, 3), rgb)
    with pytest.raises(rtb.RtbError):
        rtb.write_frame(str(tmp_path / "no_such_dir" / "f.ppm"), frame, W, H)


def test_tree_cache_round_trip(rtb, tmp_path):
    pts = rtb.geodesic_mesh(6)
    a = rtb.Trixel(pts, require_device=False)
    a.create_kd(where=1)
    path = tmp_path / "mesh.kd"
    a.save_tree(str(path))
    b = rtb.Trixel(pts, require_device=False)
    b.load_tree(str(path))
    ta, tb = a.tree(), b.tree()
    for k in ta:
        assert np.array_equal(np.asarray(ta[k]).view(np.uint8), np.asarray(tb[k]).view(np.uint8)), k
    # a cache of another mesh, a truncated file and a damaged child index are refused
    other = rtb.Trixel(pts[::-1].copy(), require_device=False)
    with pytest.raises(rtb.RtbError):
        other.load_tree(str(path))
    blob = path.read_bytes()
    (tmp_path / "short.kd").write_bytes(blob[:len(blob) // 2])
    with pytest.raises(rtb.RtbError):
        b.load_tree(str(tmp_path / "short.kd"))
    bad = bytearray(blob)
    N = 2 * len(pts) - 1
    assert blob[:8] == b"RTBKD2\0\0"
    off = 40 + 4 * 6 * N  # first entry of left[] (header: magic, num_tri, num_nodes, point hash, payload hash)
    bad[off:off + 4] = (N + 5).to_bytes(4, "little")
    (tmp_path / "bad.kd").write_bytes(bytes(bad))
    with pytest.raises(rtb.RtbError, match="payload hash"):
        b.load_tree(str(tmp_path / "bad.kd"))
    a.close(); b.close(); other.close()


def _tree_file_v1(pts, left, tri, cut):
    """A version-1 cache file (no payload hash) with the given structure: what a crafted file could hold."""
    import struct
    N = len(left)
    h = 0xcbf29ce484222325
    for byte in np.ascontiguousarray(pts, np.float32).tobytes():
        h = ((h ^ byte) * 0x100000001b3) & 0xffffffffffffffff
    return (b"RTBKD1\0\0" + struct.pack("<QQQ", len(pts), N, h) + np.zeros(6 * N, np.float32).tobytes() + np.asarray(left, np.int32).tobytes() +
            np.asarray(tri, np.int32).tobytes() + np.zeros(2 * N, np.float32).tobytes() + np.asarray(cut, np.uint8).tobytes())


def test_tree_cache_rejects_structures_the_kernels_cannot_walk(rtb, tmp_path):
    """a structurally valid chain deeper than the traversal stack, a node with two parents and a triangle named by
    two leaves must all be refused (the render kernels push without a bound check)."""
    n = 48
    pts = rtb.geodesic_mesh(2)[:n].copy()
    m = rtb.Trixel(pts, require_device=False)
    N = 2 * n - 1
    # (1) a left-leaning chain: node 2k has children 2k+1 (leaf) and 2k+2 -> depth 47 > 38, otherwise a perfectly valid tree
    left = [-1] * N; tri = [-1] * N
    t = 0
    for i in range(0, N - 1, 2):
        left[i] = i + 1
        tri[i + 1] = t; t += 1
    tri[N - 1] = t
    path = tmp_path / "deep.kd"
    path.write_bytes(_tree_file_v1(pts, left, tri, [0] * N))
    with pytest.raises(rtb.RtbError, match="deeper|damaged"):
        m.load_tree(str(path))
    # the same chain cut off at an allowed depth loads (the validator is not simply refusing version-1 files)
    n2 = 20
    pts2 = pts[:n2].copy(); m2 = rtb.Trixel(pts2, require_device=False); N2 = 2 * n2 - 1
    left2 = [-1] * N2; tri2 = [-1] * N2; t = 0
    for i in range(0, N2 - 1, 2):
        left2[i] = i + 1; tri2[i + 1] = t; t += 1
    tri2[N2 - 1] = t
    ok = tmp_path / "ok.kd"
    ok.write_bytes(_tree_file_v1(pts2, left2, tri2, [0] * N2))
    m2.load_tree(str(ok))
    # (2) two parents: node 0 and node 2 both name node 3 as their left child
    two = list(left2); two[2] = 1
    (tmp_path / "two.kd").write_bytes(_tree_file_v1(pts2, two, tri2, [0] * N2))
    with pytest.raises(rtb.RtbError):
        m2.load_tree(str(tmp_path / "two.kd"))
    # (3) one triangle named by two leaves (and one by none)
    dup = list(tri2); dup[N2 - 1] = 0
    (tmp_path / "dup.kd").write_bytes(_tree_file_v1(pts2, left2, dup, [0] * N2))
    with pytest.raises(rtb.RtbError):
        m2.load_tree(str(tmp_path / "dup.kd"))
    m.close(); m2.close()


def test_set_matrix_is_render_only_state(rtb):
    """rtb_object_set_matrix replaces the rows but not the recurrence's quaternion / faces; transform calls are
    refused until add_object restarts the recurrence (no GPU: handles are created without device memory)."""
    import ctypes
    lib = rtb.lib
    o = ctypes.c_void_p()
    pts = rtb.geodesic_mesh(2)
    m = rtb.Trixel(pts, require_device=False)
    assert lib.rtb_object_create(m.h, ctypes.byref(o)) == 0
    m12 = (ctypes.c_float * 12)(*range(12))
    assert lib.rtb_object_set_matrix(o, m12) == 5  # RTB_ERR_STATE: not added to a camera
    lib.rtb_object_destroy(o)
    m.close()


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (the arm the driver runs beside ours) needs no GPU: one bounded sample of the bunny
    through the reference's own host-compiled kernels (or the C port), one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from common import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "bunny_960x540", "--steps", "1",
                        "--warmup", "0", "--ref-frames", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "bunny_960x540"
    assert line["product_library_loaded"] is False      # the reference arm must not touch librtb.so


def test_reference_arm_standin_mesh_is_the_products_standin_mesh(rtb):
    """oracle/standin.py (numpy; what `bench.py --impl reference` and the cpu_baseline leg render when the Stanford meshes
    are absent) against rtb_mesh_geodesic (what our arm renders): the two arms must see the same triangles, bit for bit."""
    from oracle import standin
    for kw in (dict(nu=1), dict(nu=2), dict(nu=13), dict(nu=59), dict(nu=7, radius=0.09, center=(0.2, -0.1, 0.3), displacement=0.08, seed=77)):
        a = rtb.geodesic_mesh(**kw)
        b = standin.geodesic_mesh(**kw)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), kw
    big_a, big_b = rtb.geodesic_mesh(209), standin.geodesic_mesh(209)     # the dragon-sized stand-in of the headline bench
    assert np.array_equal(big_a.view(np.uint32), big_b.view(np.uint32))


def test_host_fill_machinery_without_a_gpu(rtb):
    """The sweep's background pre-fill (rtb::fill_words: streaming stores on a persistent thread pool) through its probe:
    repeated jobs with changing thread counts, unaligned heads and tails, contents checked inside the library."""
    for threads in (1, 3, 8, 2, 0, 5):
        gbs = rtb.measure_host_fill_bandwidth((1 << 20) + 4096 * (threads + 1), threads)
        assert gbs > 0.0
    with pytest.raises(rtb.RtbError):
        rtb.measure_host_fill_bandwidth(1000, 1)
