# frames mode (whole-sweep steps) for N = $1 -> gpurun_out/r1g_*_frames_n$1.json
N=$1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29801 bench.py --gpus $N > gpurun_out/r1g_dragon_orbit_frames_n$N.json 2> gpurun_out/r1g_dragon_orbit_frames_n$N.err
$R --master-port 29802 bench.py --gpus $N --workload happy_orbit_3840x2160 > gpurun_out/r1g_happy_frames_n$N.json 2> gpurun_out/r1g_happy_frames_n$N.err
for f in gpurun_out/r1g_*_frames_n$N.json; do python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), 'Mrays/s', round(d['config']['fps']), 'FPS; e2e fps', round(d['e2e']['fps']), d['clocks'])" $f; done
tail -c 300 gpurun_out/r1g_*_frames_n$N.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -5
