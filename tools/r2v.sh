set -x
timeout 900 python tools/push_tune.py unit_shift=7,8 t_active=4,8,12 t_leaf=4,8 2>&1 | tail -14
timeout 900 python tools/push_tune.py unit_shift=7 t_active=12 t_leaf=8 tail5=2,5,10,20 2>&1 | tail -5
RTB_TUNE_WORLD=8 RTB_TUNE_FRAMES=192 timeout 900 python tools/push_tune.py unit_shift=7,8 t_active=4,12 t_leaf=4,8 2>&1 | tail -10
