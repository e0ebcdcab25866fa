set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu"
$B > gpurun_out/r1c_plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv $B > gpurun_out/r1c_ncu_launch.log 2>&1
$B > gpurun_out/r1c_plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r1c_render_default -f $B > gpurun_out/r1c_ncu_full.log 2>&1
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --workload dragon_closeup_960x540"
$C > gpurun_out/r1c_plain_closeup.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r1c_render_closeup -f $C > gpurun_out/r1c_ncu_closeup.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ls -la gpurun_out/ | tail -12
