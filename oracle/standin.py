"""TEST INFRASTRUCTURE ONLY -- the labelled stand-in mesh of BASELINE.json configs[2..4] (displaced geodesic icosphere)
generated WITHOUT the product, for `bench.py --impl reference` and the cpu_baseline leg: the reference arm must not load
librtb.so.  A numpy restatement of rtb::make_geodesic (cpp_cuda_raytracer_dev_b200/csrc/rtb_host.cpp) in the same
operation order, all arithmetic in float64 with one rounding to float32 at the end; tests/test_host_cpu.py asserts that
it is bit-identical to rtb_mesh_geodesic.  (The Stanford meshes themselves are absent from the reference checkout,
/root/reference/.MISSING_LARGE_BLOBS.)"""
import numpy as np

_T = (1.0 + np.sqrt(5.0)) / 2.0
_ICO = np.array([[-1, _T, 0], [1, _T, 0], [-1, -_T, 0], [1, -_T, 0], [0, -1, _T], [0, 1, _T],
                 [0, -1, -_T], [0, 1, -_T], [_T, 0, -1], [_T, 0, 1], [-_T, 0, -1], [-_T, 0, 1]], np.float64)
_FACES = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                   [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                   [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)


def _rotl13(h):
    return (h << np.uint32(13)) | (h >> np.uint32(19))


def _hash3(x, y, z, seed):
    with np.errstate(over="ignore"):
        h = np.uint32(seed) * np.uint32(0x9E3779B1)
        h = h ^ (x.astype(np.uint32) * np.uint32(0x85EBCA77)); h = _rotl13(h); h = h * np.uint32(0xC2B2AE3D)
        h = h ^ (y.astype(np.uint32) * np.uint32(0x27D4EB2F)); h = _rotl13(h); h = h * np.uint32(0x165667B1)
        h = h ^ (z.astype(np.uint32) * np.uint32(0x9E3779B1)); h = _rotl13(h); h = h * np.uint32(0x85EBCA77)
        h = h ^ (h >> np.uint32(15)); h = h * np.uint32(0x2C1B3C6D); h = h ^ (h >> np.uint32(12)); h = h * np.uint32(0x297A2D39)
        h = h ^ (h >> np.uint32(15))
    return h


def _value_noise(x, y, z, seed):
    fx, fy, fz = np.floor(x), np.floor(y), np.floor(z)
    ix, iy, iz = fx.astype(np.int32), fy.astype(np.int32), fz.astype(np.int32)
    tx, ty, tz = x - fx, y - fy, z - fz
    tx = tx * tx * (3.0 - 2.0 * tx); ty = ty * ty * (3.0 - 2.0 * ty); tz = tz * tz * (3.0 - 2.0 * tz)
    acc = np.zeros_like(x)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        wgt = (tx if dx else 1.0 - tx) * (ty if dy else 1.0 - ty) * (tz if dz else 1.0 - tz)
        val = (_hash3(ix + np.int32(dx), iy + np.int32(dy), iz + np.int32(dz), seed) >> np.uint32(8)).astype(np.float64) * (1.0 / 8388607.5) - 1.0
        acc = acc + wgt * val
    return acc


def _vertices(face, i, j, nu, radius, center, displacement, seed):
    a = (nu - i - j).astype(np.float64) / nu
    b = i.astype(np.float64) / nu
    c = j.astype(np.float64) / nu
    d = [a * _ICO[face[0]][k] + b * _ICO[face[1]][k] + c * _ICO[face[2]][k] for k in range(3)]
    inv = 1.0 / np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    d = [np.floor((v * inv) * 1048576.0 + 0.5) / 1048576.0 for v in d]
    nz = 0.65 * _value_noise(d[0] * 3.0 + 7.3, d[1] * 3.0 + 1.9, d[2] * 3.0 + 4.1, seed) + \
        0.35 * _value_noise(d[0] * 11.0 + 0.5, d[1] * 11.0 + 8.2, d[2] * 11.0 + 2.7, seed ^ 0x5bd1e995)
    r = float(np.float32(radius)) * (1.0 + float(np.float32(displacement)) * nz)
    return np.stack([(float(np.float32(center[k])) + r * d[k]).astype(np.float32) for k in range(3)], axis=-1)


def geodesic_mesh(nu, radius=0.08, center=(0.0, 0.1, 0.0), displacement=0.05, seed=1234):
    """(20*nu*nu, 9) float32 triangle soup, triangle order and vertex order of rtb::make_geodesic."""
    # per face: for i, for j (i + j < nu): the "up" triangle, then (if i + j < nu - 1) the "down" triangle
    ii, jj, kind = [], [], []
    for i in range(nu):
        j = np.arange(nu - i)
        has_down = (i + j) < nu - 1
        # interleave up / down in emission order
        order_i = np.repeat(i, len(j) + int(has_down.sum()))
        ks = np.zeros(len(order_i), np.int64)
        js = np.zeros(len(order_i), np.int64)
        pos = np.arange(len(j)) + np.concatenate([[0], np.cumsum(has_down[:-1])]) if len(j) else np.zeros(0, np.int64)
        js[pos] = j
        down_pos = pos[has_down] + 1
        js[down_pos] = j[has_down]
        ks[down_pos] = 1
        ii.append(order_i); jj.append(js); kind.append(ks)
    ii = np.concatenate(ii); jj = np.concatenate(jj); kind = np.concatenate(kind)
    up = kind == 0
    # vertex (i, j) offsets of the three corners: up = (i,j),(i+1,j),(i,j+1); down = (i+1,j),(i+1,j+1),(i,j+1)
    ci = np.stack([np.where(up, ii, ii + 1), ii + 1, ii], axis=1)
    cj = np.stack([jj, np.where(up, jj, jj + 1), jj + 1], axis=1)
    out = np.empty((20, len(ii), 3, 3), np.float32)
    for f in range(20):
        out[f] = _vertices(_FACES[f], ci, cj, nu, radius, center, displacement, seed)
    return out.reshape(-1, 9)
