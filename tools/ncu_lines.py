"""Summarise an ncu report per CUDA source line: instructions executed, SIMT efficiency, stall samples.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-index] [top-N]
"""
import csv, subprocess, sys, io
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split per kernel
kernels, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = []
        kernels.append(cur)
    if cur is not None:
        cur.append(r)
# one block per (source file, kernel); merge all blocks of the kernel launch `which`
names = []
for b in kernels:
    fn = next((r[1] for r in b if r and r[0] == "Function Name"), "?")
    if fn not in names:
        names.append(fn)
sel = names[which] if which < len(names) else names[0]
k = [r for b in kernels if next((r[1] for r in b if r and r[0] == "Function Name"), "?") == sel for r in
     ([["FILE", next((r[1] for r in b if r and r[0] == "File Path"), "?")]] + b)]
hdr = next(r for r in k if r and r[0] == "Line No")
ix = {name: i for i, name in enumerate(hdr)}
lines = []
tot_inst = tot_thr = tot_samp = 0
curfile = ""
for r in k:
    if r and r[0] == "FILE":
        curfile = r[1].split("/")[-1].replace("rtb_", "").replace(".cuh", "")
        continue
    if len(r) != len(hdr) or r[0] in ("Line No", ""):
        continue
    r = list(r); r[0] = curfile[:7] + ":" + r[0]
    try:
        inst = int(r[ix["Instructions Executed"]]); thr = int(r[ix["Thread Instructions Executed"]]); samp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    st = {n[6:]: int(r[i]) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n and r[i].isdigit() and int(r[i])}
    lines.append((samp, inst, thr, r[0], r[1], st))
    tot_inst += inst; tot_thr += thr; tot_samp += samp
print("kernel %d: warp-inst %.3fG thread-inst %.3fG avg-threads %.1f samples %d" % (which, tot_inst / 1e9, tot_thr / 1e9, tot_thr / max(tot_inst, 1), tot_samp))
for samp, inst, thr, ln, src, st in sorted(lines, reverse=True)[:top]:
    top_st = ",".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% samp %5.1f%% inst  thr/inst %4.1f  L%-12s %-70s %s" % (100.0 * samp / tot_samp, 100.0 * inst / tot_inst, thr / max(inst, 1), ln, src.strip()[:70], top_st))
