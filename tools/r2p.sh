# round 2: L2 persisting window once more, now that the frames of a launch are worked through in view order
set -x
export RTB_TUNE_FRAMES=600
timeout 900 python tools/tune.py L2_WINDOW=0,1,2 2>&1 | tail -5
timeout 900 python tools/tune.py L2_WINDOW=2 L2_CARVE_MB=40,64,79 2>&1 | tail -4
timeout 900 python tools/tune.py L2_WINDOW=1 L2_CARVE_MB=24,40 2>&1 | tail -3
