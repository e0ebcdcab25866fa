"""Build several variants of librtb.so side by side (development aid): `name=-DFLAG=1,-DOTHER=2 ...` -> build/variants/librtb_<name>.so.
Select one at run time with RTB_LIB=<path> (cpp_cuda_raytracer_dev_b200/__init__.py)."""
import importlib.util, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("rtb_build", os.path.join(ROOT, "cpp_cuda_raytracer_dev_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
procs = []
for arg in sys.argv[1:]:
    name, _, flags = arg.partition("=")
    out = os.path.join(out_dir, "librtb_%s.so" % name)
    cmd = [b.nvcc_path(), "-ccbin", "/usr/bin/g++"] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
    procs.append((name, out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, out, p in procs:
    text = p.communicate()[0]
    regs = [l for l in text.splitlines() if "render_stream_kernelILb1ELb0ELb0" in l or "Used" in l]
    print(name, "rc", p.returncode, out)
    keep = False
    for l in text.splitlines():
        if "Compiling entry function" in l:
            keep = "render_stream_kernel" in l
            if keep: print("   ", l.split("'")[1][:60])
        elif keep and ("Used" in l or "spill" in l):
            print("      ", l.strip())
    if p.returncode != 0:
        print(text[-3000:])
