# round 2: single-frame knobs (prefetch, steal spin, t_active of inline launches), sweep into pinned buffers, full GPU tests
set -x
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -8
for env in "" "RTB_INLINE_PREFETCH=0" "RTB_T_ACTIVE_INLINE=12" "RTB_STEAL_SPIN=4" "RTB_STEAL_SPIN=16" "RTB_STEAL_SPIN=32 RTB_T_ACTIVE_INLINE=31"; do
  echo "=== ${env:-default}"
  env $env timeout 300 python tools/frame_probe.py 2>&1 | grep -v "unit_shift\|t_active\|no root\|same loop\|D2H of\|transform  \|lib:" | tail -6
done
RTB_LIB=build/variants/librtb_wlog.so timeout 300 python tools/warp_log.py 1 2>&1 | tail -16
timeout 900 python bench.py --steps 5 --no-other-workloads > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'e2e fps',round(d['e2e']['fps']),'d2h_gbs',d['e2e'].get('d2h_gbs'),'match',d['e2e']['matches_device_run'],'loop fps',round(d['frame_loop']['fps']),'seam fps',round(d['reference_classes_over_librtb']['fps']))
PY
RTB_SWEEP_DIRECT=0 timeout 900 python bench.py --steps 5 --no-other-workloads --no-cpu-baseline > gpurun_out/r2j_bench_ring.json 2> gpurun_out/r2j_bench_ring.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_ring.json')); print('ring: e2e',round(d['e2e']['value']),'fps',round(d['e2e']['fps']))"
