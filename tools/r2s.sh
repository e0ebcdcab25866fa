# round 2: re-tune of the multi-frame kernel's knobs now that the frames are view-ordered; 7 and 6 blocks per SM
set -x
export RTB_TUNE_FRAMES=600
timeout 900 python tools/tune.py T_ACTIVE=8,12,16,20 T_LEAF=4,8 2>&1 | tail -10
timeout 900 python tools/tune.py UNIT_SHIFT=6,7,8 2>&1 | tail -4
timeout 900 python tools/tune.py TAIL5=0,15,30,60 2>&1 | tail -5
for v in min7 min6; do
  echo "=== variant $v"
  RTB_LIB=build/variants/librtb_$v.so timeout 600 python tools/tune.py T_ACTIVE=12,16 2>&1 | tail -3
done
