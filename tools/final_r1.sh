# round-1 final record on one B200: tests, smoke, both bench arms, ncu launch list + full captures of the same commands
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --impl reference > gpurun_out/r1g_reference_arm.json 2> gpurun_out/r1g_reference_arm.err
python bench.py > gpurun_out/r1g_dragon_orbit_full_n1.json 2> gpurun_out/r1g_dragon_orbit_full_n1.err
python bench.py --workload dragon_closeup_960x540 --steps 5 > gpurun_out/r1g_dragon_closeup_full_n1.json 2> gpurun_out/r1g_dragon_closeup_full_n1.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu"
$B > gpurun_out/r1g_plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"render_stream|pack_|fill|compose" -c 60 --csv --log-file gpurun_out/r1g_launches.csv $B > gpurun_out/r1g_ncu_launch.log 2>&1
$B > gpurun_out/r1g_plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r1g_render_default -f $B > gpurun_out/r1g_ncu_full.log 2>&1
D="python tools/push_local_probe.py 233 3840 2160 6 0 2"
$D > gpurun_out/r1g_plain_push.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 10 -c 1 -o gpurun_out/r1g_render_push -f $D > gpurun_out/r1g_ncu_push.log 2>&1
ls -la gpurun_out | grep r1g_ | tail -20
