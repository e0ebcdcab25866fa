# round 2, N = $1 GPUs: the driver's scaling line (frames mode + the `tiles` leg: 4K split by tiles, striped peer push) and a
# dense tiles-mode line (64 % coverage), each under a timeout.  Raw JSON -> gpurun_out/r2_*_n$1.json
N=$1
R="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29811 bench.py --gpus $N --steps 5 > gpurun_out/r2_dragon_orbit_frames_n$N.json 2> gpurun_out/r2_dragon_orbit_frames_n$N.err
$R --master-port 29812 bench.py --gpus $N --workload dragon_closeup_960x540 --steps 5 --shard tiles --no-tiles-leg > gpurun_out/r2_closeup_tiles_push_n$N.json 2> gpurun_out/r2_closeup_tiles_push_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for name in ("dragon_orbit_frames", "closeup_tiles_push"):
    f = "gpurun_out/r2_%s_n%s.json" % (name, N)
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(name, "value", round(d["value"]), "Mrays/s; e2e", round(d["e2e"]["value"]), "; d2h copy-only GB/s", round(d["e2e"]["d2h_copy_only_gbs"], 1), d["clocks"])
        if d.get("tiles"):
            t = d["tiles"]
            print("  tiles leg:", round(t["value"]), "Mrays/s,", round(t["fps"]), "FPS; one GPU same frames", round(t["one_gpu_value"]), "-> efficiency", round(t["efficiency_vs_one_gpu"], 3), "equal:", t["pushed_frames_equal_single_gpu_frames"])
        print("  parallelism:", d["config"]["parallelism"][:150])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/r2_*_n$N.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -8
