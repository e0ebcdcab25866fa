# round 2 record, first pass: GPU tests, smoke, both bench arms, ncu launch list + full captures of the same commands
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --impl reference > gpurun_out/r2f_reference_arm.json 2> gpurun_out/r2f_reference_arm.err; tail -c 400 gpurun_out/r2f_reference_arm.json
timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-other-workloads"
$B > gpurun_out/r2f_plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"render_|pack_|fill|compose|scene" -c 80 --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/r2f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r2f_render_default -f $B > gpurun_out/r2f_ncu_full.log 2>&1
C="$B --workload dragon_closeup_960x540"
$C > gpurun_out/r2f_plain_closeup.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r2f_render_closeup -f $C > gpurun_out/r2f_ncu_closeup.log 2>&1
timeout 600 python tools/frame_probe.py 2>&1 | tail -24
ls -la gpurun_out | grep r2f_
