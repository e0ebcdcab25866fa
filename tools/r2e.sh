# round 2: GPU tests with the scene extension, full bench, then 2 GPUs: frames bench with the tiles leg
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'loop',d['frame_loop']['fps'])
print(json.dumps(d['scene_extension'],indent=1))
PY
