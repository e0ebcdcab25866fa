# round 2, closing record: GPU tests, smoke, both bench arms on the final code
set -x
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --impl reference > gpurun_out/r2q_reference_arm.json 2> gpurun_out/r2q_reference_arm.err; tail -c 300 gpurun_out/r2q_reference_arm.json
timeout 1500 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench.json'))
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),round(d['e2e']['fps']),'loop',{k:round(v) for k,v in d['frame_loop'].items() if k.startswith('fps')},'seam',round(d['reference_classes_over_librtb']['fps']),'refgpu',round(d['reference_gpu']['fps']),'cpu',round(d['cpu_baseline']['value'],1))
print('roofline',r['bound'],round(r['frac'],3),{k:round(v['frac'],3) for k,v in r['legs'].items()},r['traffic'],r['traffic_source'])
for k,v in d['workloads'].items(): print(k, round(v['value']), round(v['fps']),'e2e',round(v['e2e']['value']),round(v['e2e']['fps']),'loop',round(v['frame_loop_fps']),'build',round(v['tree_build_s'],4))
PY
