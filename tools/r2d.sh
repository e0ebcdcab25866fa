set -x
RTB_LIB=build/variants/librtb_wlog.so timeout 300 python tools/warp_log.py 2 2>&1 | tail -40
RTB_LIB=build/variants/librtb_wlog.so timeout 300 python tools/warp_log.py 1 7 2>&1 | tail -20
