# multi-GPU bench lines for N = $1: frames mode (dragon orbit), tiles mode with the fused peer push (4K Buddha stand-in,
# 8K synthetic 10M), and the NCCL gather path beside it at 4K.  Raw JSON -> gpurun_out/r1d_*_n$1.json
N=$1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29601 bench.py --gpus $N > gpurun_out/r1d_dragon_orbit_frames_n$N.json 2> gpurun_out/r1d_dragon_orbit_frames_n$N.err
$R --master-port 29602 bench.py --gpus $N --workload happy_orbit_3840x2160 --steps 5 > gpurun_out/r1d_happy_frames_n$N.json 2> gpurun_out/r1d_happy_frames_n$N.err
$R --master-port 29603 bench.py --gpus $N --workload happy_orbit_3840x2160 --steps 5 --shard tiles > gpurun_out/r1d_happy_tiles_push_n$N.json 2> gpurun_out/r1d_happy_tiles_push_n$N.err
$R --master-port 29604 bench.py --gpus $N --workload happy_orbit_3840x2160 --steps 5 --shard tiles --exchange nccl > gpurun_out/r1d_happy_tiles_nccl_n$N.json 2> gpurun_out/r1d_happy_tiles_nccl_n$N.err
$R --master-port 29605 bench.py --gpus $N --workload synthetic10m_7680x4320 --steps 5 --shard tiles > gpurun_out/r1d_synthetic10m_tiles_push_n$N.json 2> gpurun_out/r1d_synthetic10m_tiles_push_n$N.err
for f in gpurun_out/r1d_*_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), 'Mrays/s', round(d['config']['fps']), 'FPS; e2e fps', round(d['e2e']['fps']), d['config']['parallelism'][:60])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -c 400 gpurun_out/r1d_*_n$N.err | grep -v "^$" | tail -12
