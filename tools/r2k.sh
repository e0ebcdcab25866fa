# round 2: frame slots + lookahead -- full GPU tests, single-frame loop with and without frames ahead, seam loop
set -x
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
for look in 2 1 0; do
  echo "=== lookahead=$look"
  RTB_LOOKAHEAD=$look timeout 300 python tools/frame_probe.py 2>&1 | grep "full loop\|render + color\|same loop\|single-frame kernel, orbit"
done
timeout 900 python bench.py --steps 5 --no-other-workloads > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'loop',{k:round(v) for k,v in d['frame_loop'].items() if k.startswith('fps')},'seam fps',round(d['reference_classes_over_librtb']['fps']),'ref gpu fps',round(d['reference_gpu']['fps']))
PY
