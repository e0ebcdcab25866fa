# round 2, second GPU call: restructured API (tests), loop-control / stack variants, single-frame anatomy
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()"
export RTB_TUNE_FRAMES=600
for v in base morton redux packed both; do
  echo "=== variant $v"
  RTB_LIB=build/variants/librtb_$v.so timeout 600 python tools/tune.py T_ACTIVE=12,16 2>&1 | tail -5
done
echo "=== frame anatomy"
timeout 600 python tools/frame_probe.py 2>&1 | tail -22
echo "=== bench (no extras)"
timeout 900 python bench.py --steps 5 --no-other-workloads > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 3000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
