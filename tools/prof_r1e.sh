# ncu evidence for the round-1 final kernels.  Every profiled command first exits 0 without ncu.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu"
$B > gpurun_out/r1e_plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"render_stream|pack_|fill|Fill|compose" -c 200 --csv --log-file gpurun_out/r1e_launches.csv $B > gpurun_out/r1e_ncu_launch.log 2>&1
$B > gpurun_out/r1e_plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r1e_render_default -f $B > gpurun_out/r1e_ncu_full.log 2>&1
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --workload dragon_closeup_960x540"
$C > gpurun_out/r1e_plain_closeup.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -o gpurun_out/r1e_render_closeup -f $C > gpurun_out/r1e_ncu_closeup.log 2>&1
D="python tools/push_local_probe.py 233 3840 2160 6 0 2"
$D > gpurun_out/r1e_plain_push.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"render_stream_kernel<1, 0, 1>" -s 2 -c 1 -o gpurun_out/r1e_render_push -f $D > gpurun_out/r1e_ncu_push.log 2>&1
E="python tools/build_probe.py 209"
$E > gpurun_out/r1e_plain_build.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1e_build_launches.csv $E > gpurun_out/r1e_ncu_build.log 2>&1
ls -la gpurun_out | grep r1e
