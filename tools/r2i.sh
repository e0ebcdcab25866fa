# round 2: per-warp anatomy of a single frame with and without ray sharing; builder after the side-byte change
set -x
for v in wlog wlog_nosteal; do
  echo "=== $v"
  RTB_LIB=build/variants/librtb_$v.so timeout 300 python tools/warp_log.py 2 2>&1 | tail -32
done
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "tree_build or headless or icosphere" 2>&1 | tail -5
timeout 300 python tools/build_probe.py 209 2>&1 | tail -6
timeout 300 python tools/build_probe.py 707 2>&1 | tail -6
