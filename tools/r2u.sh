# round 2 record, final code after the knob re-tune: GPU tests, smoke, both bench arms, ncu launch list + full captures of the same commands
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --impl reference > gpurun_out/r2u_reference_arm.json 2> gpurun_out/r2u_reference_arm.err; tail -c 400 gpurun_out/r2u_reference_arm.json
timeout 1500 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2u_bench.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-other-workloads"
$B > gpurun_out/r2u_plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"render_|pack_|fill|compose|scene" -c 80 --csv --log-file gpurun_out/r2u_launches.csv $B > gpurun_out/r2u_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r2u_render_default -f $B > gpurun_out/r2u_ncu_full.log 2>&1
C="$B --workload dragon_closeup_960x540"
$C > gpurun_out/r2u_plain_closeup.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r2u_render_closeup -f $C > gpurun_out/r2u_ncu_closeup.log 2>&1
timeout 600 python tools/frame_probe.py 2>&1 | tail -24
ls -la gpurun_out | grep r2u_
python tools/one_frame.py 40 > gpurun_out/r2u_one_frame_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 30 -c 1 -o gpurun_out/r2u_one_frame -f python tools/one_frame.py 40 > gpurun_out/r2u_one_frame_ncu.log 2>&1
tail -2 gpurun_out/r2u_one_frame_plain.log
E="python tools/build_probe.py 209"
$E > gpurun_out/r2u_plain_build.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2u_build_launches.csv $E > gpurun_out/r2u_ncu_build.log 2>&1
tail -4 gpurun_out/r2u_plain_build.log
ls -la gpurun_out | grep r2u_ | tail -20
RTB_TUNE_FRAMES=600 timeout 900 python tools/tune.py T_ACTIVE=3,4,5,6 T_LEAF=3,4,5 2>&1 | tail -13
