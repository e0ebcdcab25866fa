# round 2, N = $1 GPUs: the dense (64 % coverage) 960x540 configuration split by tiles, striped fused push, final code
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus $N --workload dragon_closeup_960x540 --steps 5 --shard tiles --no-tiles-leg > gpurun_out/r2d_closeup_tiles_push_n$N.json 2> gpurun_out/r2d_closeup_tiles_push_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2d_closeup_tiles_push_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("closeup tiles push: value", round(d["value"]), "Mrays/s", round(d["config"]["fps"]), "FPS, ms/step", round(d["ms_per_step"], 3), d["config"]["parallelism"][-60:])
PY
tail -c 300 gpurun_out/r2d_closeup_tiles_push_n$N.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
