# round 2, first GPU call: kernel memory variants side by side (tools/build_variants.py) + single-frame loop costs
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv
export RTB_TUNE_FRAMES=${RTB_TUNE_FRAMES:-600}
for v in base s4 s6 s8 hint s6hint; do
  echo "=== variant $v"
  RTB_LIB=build/variants/librtb_$v.so timeout 600 python tools/tune.py L2_WINDOW=scene,nodes L2_PERSIST=1 2>&1 | tail -8
done
echo "=== persist off"
RTB_LIB=build/variants/librtb_base.so timeout 600 python tools/tune.py L2_PERSIST=0 2>&1 | tail -3
RTB_LIB=build/variants/librtb_s6.so timeout 600 python tools/tune.py L2_PERSIST=0 2>&1 | tail -3
echo "=== frame loop (base)"
RTB_LIB=build/variants/librtb_base.so timeout 300 python tools/loop_probe.py 2>&1 | tail -8
