# round 2: sweep into pinned buffers with growing chunks on two alternating streams; lookahead back-off; full GPU tests
set -x
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -6
for mb in 512 256 1024; do
  echo "=== sweep_chunk_mb=$mb"
  RTB_SWEEP_CHUNK_MB=$mb timeout 900 python bench.py --steps 6 --no-other-workloads --no-cpu-baseline > gpurun_out/r2n_bench_$mb.json 2> gpurun_out/r2n_bench_$mb.err; tail -2 gpurun_out/r2n_bench_$mb.err
  python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_$mb.json')); e=d['e2e']; print('value',round(d['value']),'e2e',round(e['value']),'fps',round(e['fps']),'GB/s',round(e['d2h_gbs'],1),'fill',round(e['host_fill_gbs'],1),'match',e['matches_device_run'],'loop',round(d['frame_loop']['fps']))"
done
for w in dragon_closeup_960x540 happy_orbit_3840x2160; do
  timeout 900 python bench.py --steps 3 --workload $w --no-other-workloads --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$w value',round(d['value']),'e2e',round(e['value']),'fps',round(e['fps']),'match',e['matches_device_run'],'loop',round(d['frame_loop']['fps']))"
done
