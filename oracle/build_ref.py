#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- build oracle/_ref/libref_emu.so from the reference's own sources.

The reference (ams3878/cpp_cuda_raytracer_dev) is a Win32 + CUDA 11.6 Visual Studio project with no
CPU render path.  Its ray-cast path is nevertheless plain C++ inside `__global__` functions, so this
recipe compiles the reference's OWN files for the host:

  * sources are read IN PLACE from $RTB_REFERENCE_DIR (default /root/reference/TEST_Dungeonrun),
    patched in memory and streamed to g++ on stdin as ONE translation unit -- no reference source
    is ever written into this repository; the only outputs are under oracle/_ref/ (git-ignored);
  * oracle/shim/ supplies a host emulation of the CUDA runtime calls used (cudaMalloc -> calloc,
    a launch -> loop nest over blocks/threads, blocks spread over host cores with OpenMP) and the
    few Win32 names the headers mention;
  * in-memory patches (portability only, no arithmetic is touched):
      1. Vector.h   : `using VEC3<T>::x/y/z;` inside VEC4 (MSVC permissive dependent-base lookup)
      2. vector.cpp : `template<>` on the three explicit member specialisations
      3. read_ply.cpp: `a[i].id = off + i++;` split in two statements (unsequenced in MSVC's favour)
      4. *.cu       : `k << < G, B >> > (args)` -> `EMU_LAUNCH(k, G, B, args)`
      5. vector.cpp : `union { float x; s64 i; }` -> s32 (upper half uninitialised in the original;
                      MSVC leaves it zero, which is the 32-bit behaviour)
  * flags: -O2 -ffp-contract=off (the north star's "fp32, FMA contraction off" oracle), SSE2 floats.

The driver that replays WinMain.cpp's call sequence headlessly is oracle/ref_driver.cpp (ours).
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RTB_REFERENCE_DIR", "/root/reference/TEST_Dungeonrun")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libref_emu.so")

ORDER = ["vector.cpp", "Quaternion.cpp", "Color.cpp", "Input.cpp", "Object.cpp", "Camera.cpp", "read_ply.cpp",
         "Quaternion.cu", "Camera.cu", "Trixel.cu"]


def _match_paren(text, start):
    depth = 0
    for k in range(start, len(text)):
        if text[k] == "(":
            depth += 1
        elif text[k] == ")":
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced parentheses")


def rewrite_launches(text):
    """`name << < G, B >> > (args)` -> `EMU_LAUNCH(name, (G), (B), args)`."""
    out = []
    pos = 0
    pat = re.compile(r"(\w+)\s*<<\s*<")
    while True:
        m = pat.search(text, pos)
        if not m:
            out.append(text[pos:])
            break
        close = re.compile(r">>\s*>").search(text, m.end())
        cfg = text[m.end():close.start()]
        depth, split = 0, None
        for k, ch in enumerate(cfg):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 0:
                split = k
                break
        grid, block = cfg[:split].strip(), cfg[split + 1:].strip()
        lp = text.index("(", close.end())
        rp = _match_paren(text, lp)
        args = text[lp + 1:rp]
        out.append(text[pos:m.start()])
        out.append("EMU_LAUNCH(%s, (%s), (%s), %s)" % (m.group(1), grid, block, args))
        pos = rp + 1
    return "".join(out)


def patched(name, cuda=False):
    with open(os.path.join(REF, name), "r", encoding="utf-8-sig") as f:
        text = f.read()
    if name == "Vector.h":
        anchor = "union { T w; T t; T dt; T d; };"
        assert anchor in text
        text = text.replace(anchor, "using VEC3<T>::x; using VEC3<T>::y; using VEC3<T>::z; " + anchor, 1)
    if name == "vector.cpp":
        for sig in ("T_fp VEC3<T_fp>::dot(", "void VEC4<T_fp>::cross(", "void VEC4<T_fp>::rotate("):
            assert sig in text
            text = text.replace(sig, "template<> " + sig, 1)
        assert "union { float x; s64 i; } u;" in text
        text = text.replace("union { float x; s64 i; } u;", "union { float x; s32 i; } u;", 1)
    if name == "read_ply.cpp":
        n = text.count("= triangle_index_offset + leaf_index++;")
        assert n == 3
        text = text.replace("= triangle_index_offset + leaf_index++;", "= triangle_index_offset + leaf_index; leaf_index++;")
    if name.endswith(".cu") and not cuda:
        text = rewrite_launches(text)
    return '#line 1 "%s"\n%s\n' % (os.path.join(REF, name), text)


def build(force=False):
    if not os.path.isdir(REF):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(REF, n) for n in ORDER + ["Vector.h"]] + [os.path.join(HERE, "ref_driver.cpp"), __file__,
                                                                  os.path.join(HERE, "shim", "cuda_runtime.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        copy_assets()
        return OUT
    unit = [patched("Vector.h")] + [patched(n) for n in ORDER]
    with open(os.path.join(HERE, "ref_driver.cpp")) as f:
        unit.append('#line 1 "%s"\n%s\n' % (os.path.join(HERE, "ref_driver.cpp"), f.read()))
    cmd = ["/usr/bin/g++", "-x", "c++", "-", "-std=c++17", "-O2", "-ffp-contract=off", "-fpermissive", "-w", "-fopenmp", "-fPIC",
           "-shared", "-I", os.path.join(HERE, "shim"), "-I", REF, "-o", OUT]
    r = subprocess.run(cmd, input="".join(unit).encode(), cwd=OUT_DIR)
    if r.returncode != 0:
        raise RuntimeError("reference emulation build failed")
    copy_assets()
    return OUT


CUDA_VARIANTS = {
    # name -> extra nvcc flags.  "fmad" is the reference project's own code generation (TEST_Dungeonrun.vcxproj has no
    # --fmad switch, so nvcc's default contraction applies): the kernel to beat.  "nofmad" is the same code with
    # contraction off, i.e. the north star's oracle contract executed by the reference's own kernels on the GPU.
    "fmad": [],
    "nofmad": ["-fmad=false"],
}


def cuda_lib_path(variant):
    return os.path.join(OUT_DIR, "libref_cuda_%s.so" % variant)


def build_cuda(force=False):
    """The reference's own .cu/.cpp files compiled by nvcc for sm_100a (real kernels, real launches).

    Same in-memory patches as the host build minus the launch rewrite (#4).  nvcc cannot read a translation unit
    from stdin, so the patched unit is written to a temporary directory outside the repository, compiled, and the
    directory removed; only the .so files land in oracle/_ref/."""
    import shutil
    import tempfile
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isdir(REF) or not os.path.exists(nvcc):
        return []
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(REF, n) for n in ORDER + ["Vector.h"]] + [os.path.join(HERE, "ref_driver.cpp"), __file__]
    outs = [cuda_lib_path(v) for v in CUDA_VARIANTS]
    if not force and all(os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in deps) for o in outs):
        return outs
    unit = ["#define RTB_REF_CUDA 1\n", patched("Vector.h", cuda=True)] + [patched(n, cuda=True) for n in ORDER]
    with open(os.path.join(HERE, "ref_driver.cpp")) as f:
        unit.append('#line 1 "%s"\n%s\n' % (os.path.join(HERE, "ref_driver.cpp"), f.read()))
    with tempfile.TemporaryDirectory(prefix="rtb_refcuda_") as tmp:
        src = os.path.join(tmp, "ref_unit.cu")
        with open(src, "w") as f:
            f.write("".join(unit))
        for variant, extra in CUDA_VARIANTS.items():
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-lineinfo", "-w", "-shared", "-cudart", "static",
                   "-Xcompiler", "-fPIC,-fpermissive,-w,-O2,-ffp-contract=off", "-I", os.path.join(HERE, "shim_win"), "-I", REF,
                   "-o", cuda_lib_path(variant), src] + extra
            r = subprocess.run(cmd, cwd=tmp)
            if r.returncode != 0:
                raise RuntimeError("reference CUDA build (%s) failed" % variant)
    return outs


def seam_lib_path():
    return os.path.join(OUT_DIR, "libref_seam.so")


def build_seam(force=False):
    """The drop-in itself: the reference's HOST sources (Camera.cpp, Object.cpp, Quaternion.cpp, Input.cpp, Color.cpp,
    vector.cpp, read_ply.cpp, Trixel.h -- read in place, same portability patches) compiled WITHOUT its three .cu files;
    integration/rtb_seam.cpp supplies the seam functions on top of librtb.so.  The host classes' own cudaMalloc /
    cudaMemcpy calls (dead weight once the seam is in) go to the host shim.  Needs librtb.so."""
    root = os.path.dirname(HERE)
    lib_dir = os.path.join(root, "cpp_cuda_raytracer_dev_b200")
    seam = os.path.join(root, "integration", "rtb_seam.cpp")
    if not os.path.isdir(REF) or not os.path.exists(os.path.join(lib_dir, "librtb.so")):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    host_sources = [n for n in ORDER if not n.endswith(".cu")]
    out = seam_lib_path()
    deps = [os.path.join(REF, n) for n in host_sources + ["Vector.h"]] + [os.path.join(HERE, "ref_driver.cpp"), seam, __file__,
                                                                          os.path.join(root, "include", "rtb.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    unit = ["#define RTB_REF_SEAM 1\n", patched("Vector.h")] + [patched(n) for n in host_sources]
    for path in (seam, os.path.join(HERE, "ref_driver.cpp")):
        with open(path) as f:
            unit.append('#line 1 "%s"\n%s\n' % (path, f.read()))
    cmd = ["/usr/bin/g++", "-x", "c++", "-", "-std=c++17", "-O2", "-ffp-contract=off", "-fpermissive", "-w", "-fopenmp", "-fPIC", "-shared",
           "-I", os.path.join(HERE, "shim"), "-I", REF, "-I", os.path.join(root, "include"), "-o", out,
           "-L", lib_dir, "-lrtb", "-Wl,-rpath,$ORIGIN/../../cpp_cuda_raytracer_dev_b200"]
    r = subprocess.run(cmd, input="".join(unit).encode(), cwd=OUT_DIR)
    if r.returncode != 0:
        raise RuntimeError("reference seam build failed")
    return out


def copy_assets():
    """The reference's two loadable meshes travel to the GPU box as git-ignored build outputs
    (oracle/_ref/data/): they are inputs of the parity tests, not product source."""
    import shutil
    dst = os.path.join(OUT_DIR, "data")
    os.makedirs(dst, exist_ok=True)
    for name in ("rabbit_70k.ply", "3_walls.ply"):
        src = os.path.join(REF, name)
        if os.path.exists(src) and not os.path.exists(os.path.join(dst, name)):
            shutil.copyfile(src, os.path.join(dst, name))


if __name__ == "__main__":
    path = build(force="--force" in sys.argv)
    print(path if path else "reference sources not present at %s; nothing built" % REF)
    for path in build_cuda(force="--force" in sys.argv):
        print(path)
    print(build_seam(force="--force" in sys.argv))
