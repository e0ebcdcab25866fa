# round 2: single-frame path -- ray sharing between lanes (STEAL) and frames stored straight into the host buffers
set -x
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
for v in "" nosteal spin2 spin8; do
  echo "=== variant ${v:-default}"
  for hd in 1 0; do
    echo "--- host_direct=$hd"
    if [ -n "$v" ]; then export RTB_LIB=build/variants/librtb_$v.so; else unset RTB_LIB; fi
    RTB_HOST_DIRECT=$hd timeout 300 python tools/frame_probe.py 2>&1 | grep -v "unit_shift\|t_active\|no root" | tail -12
  done
done
unset RTB_LIB
timeout 300 python tools/build_probe.py 209 2>&1 | tail -6
timeout 300 python tools/build_probe.py 707 2>&1 | tail -6
timeout 120 python -c "
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
for t in (1, 2, 4, 8, 12, 16):
    print('host fill, streaming stores, %2d threads: %.1f GB/s' % (t, rtb.measure_host_fill_bandwidth(512 << 20, t)))
"
