# tiles mode with the fused peer push (pre-filled destination) for N = $1 -> gpurun_out/r1f_*_n$1.json
N=$1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29703 bench.py --gpus $N --workload happy_orbit_3840x2160 --steps 5 --shard tiles > gpurun_out/r1f_happy_tiles_push_n$N.json 2> gpurun_out/r1f_happy_tiles_push_n$N.err
$R --master-port 29704 bench.py --gpus $N --workload dragon_closeup_960x540 --steps 5 --shard tiles > gpurun_out/r1f_closeup_tiles_push_n$N.json 2> gpurun_out/r1f_closeup_tiles_push_n$N.err
$R --master-port 29705 bench.py --gpus $N --workload synthetic10m_7680x4320 --steps 5 --shard tiles > gpurun_out/r1f_synthetic10m_tiles_push_n$N.json 2> gpurun_out/r1f_synthetic10m_tiles_push_n$N.err
for f in gpurun_out/r1f_*_n$N.json; do python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), 'Mrays/s', round(d['config']['fps']), 'FPS', d['config']['parallelism'][-40:])" $f; done
