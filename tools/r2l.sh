# round 2: ray sharing in the multi-frame kernels too (tail of a launch)?  base vs stealall, 600- and 60-frame launches, 4K
set -x
for frames in 600 60; do
  for v in "" stealall; do
    echo "=== frames=$frames variant=${v:-base}"
    if [ -n "$v" ]; then export RTB_LIB=build/variants/librtb_$v.so; else unset RTB_LIB; fi
    RTB_TUNE_FRAMES=$frames timeout 600 python tools/tune.py T_ACTIVE=12 2>&1 | tail -2
  done
done
for v in "" stealall; do
  echo "=== 4K variant=${v:-base}"
  if [ -n "$v" ]; then export RTB_LIB=build/variants/librtb_$v.so; else unset RTB_LIB; fi
  for f in 36 6; do
    timeout 600 python bench.py --workload happy_orbit_3840x2160 --frames-per-step $f --steps 6 --no-cpu-baseline --no-other-workloads 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frames/step', d['config']['frames_per_step'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'kernel', d['kernel_ms'])"
  done
done
