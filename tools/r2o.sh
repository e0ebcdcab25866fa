# round 2: sweep with a persistent fill pool and tapering chunks
set -x
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "sweep or headless or ahead" 2>&1 | tail -4
for t in 0 8 16; do
  echo "=== host_fill_threads=$t"
  RTB_HOST_FILL_THREADS=$t timeout 900 python bench.py --steps 6 --no-other-workloads --no-cpu-baseline --no-reference-gpu > gpurun_out/r2o_bench_$t.json 2> gpurun_out/r2o_bench_$t.err; tail -2 gpurun_out/r2o_bench_$t.err
  python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_$t.json')); e=d['e2e']; print('value',round(d['value']),'e2e',round(e['value']),'fps',round(e['fps']),'GB/s',round(e['d2h_gbs'],1),'fill',round(e['host_fill_gbs'],1),'match',e['matches_device_run'],'loop',round(d['frame_loop']['fps']))"
done
