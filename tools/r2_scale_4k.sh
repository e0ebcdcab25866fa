# round 2, N = $1 GPUs: the 4K configuration sharded by frames (36 frames per rank and step), no tiles leg
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus $N --steps 5 --workload happy_orbit_3840x2160 --no-tiles-leg > gpurun_out/r2_happy_frames_n$N.json 2> gpurun_out/r2_happy_frames_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2_happy_frames_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("4K frames: value", round(d["value"]), "Mrays/s", round(d["config"]["fps"]), "FPS; e2e", round(d["e2e"]["value"]), round(d["e2e"]["fps"]), "FPS", d["clocks"])
PY
tail -c 300 gpurun_out/r2_happy_frames_n$N.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -4
