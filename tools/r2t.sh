# round 2: re-tune, second pass (lower t_active / t_leaf, larger units, shorter tail segment)
set -x
export RTB_TUNE_FRAMES=600
timeout 900 python tools/tune.py UNIT_SHIFT=8 TAIL5=8 T_ACTIVE=2,4,6,8 T_LEAF=2,4,6 2>&1 | tail -14
timeout 900 python tools/tune.py UNIT_SHIFT=9 TAIL5=8 T_ACTIVE=4,8 T_LEAF=4 2>&1 | tail -3
timeout 900 python tools/tune.py UNIT_SHIFT=8 TAIL5=0,4,16 T_ACTIVE=8 T_LEAF=4 2>&1 | tail -4
RTB_TUNE_FRAMES=60 timeout 900 python tools/tune.py UNIT_SHIFT=7,8 T_ACTIVE=8,12 T_LEAF=4,8 2>&1 | tail -9
