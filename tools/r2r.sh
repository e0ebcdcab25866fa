# round 2: builder with block-split prefixes (no device-wide six-component scan)
set -x
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "tree_build or headless or icosphere or full_size" 2>&1 | tail -4
timeout 300 python tools/build_probe.py 209 2>&1 | tail -5
timeout 300 python tools/build_probe.py 707 2>&1 | tail -4
timeout 300 python tools/build_probe.py 59 2>&1 | tail -4
