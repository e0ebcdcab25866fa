# round 2, N = $1 GPUs: the driver's scaling line only (frames mode + the `tiles` leg with the pre-fill in the launch's tail)
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus $N --steps 5 > gpurun_out/r2b_dragon_orbit_frames_n$N.json 2> gpurun_out/r2b_dragon_orbit_frames_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2b_dragon_orbit_frames_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
t = d["tiles"]
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "tiles leg:", round(t["value"]), "Mrays/s,", round(t["fps"]), "FPS, ms/step", round(t["ms_per_step"], 3), "; frames_4k", round(d["frames_4k"]["value"]), round(d["frames_4k"]["per_gpu_value"]), "; one GPU", round(t["one_gpu_value"]), round(t["one_gpu_same_frames_ms"], 3), "ms -> efficiency", round(t["efficiency_vs_one_gpu"], 3), "equal:", t["pushed_frames_equal_single_gpu_frames"])
PY
tail -c 400 gpurun_out/r2b_dragon_orbit_frames_n$N.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -5
