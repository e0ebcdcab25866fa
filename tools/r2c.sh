# round 2, third GPU call: L2 carve-out sweep, single-frame anatomy with clocks + ncu, first full bench of the new bench.py
set -x
export RTB_TUNE_FRAMES=600
timeout 900 python tools/tune.py L2_CARVE_MB=8,24,40,48,56,64,79 2>&1 | tail -10
timeout 600 python tools/frame_probe.py 2>&1 | tail -24
python tools/one_frame.py 40 > gpurun_out/r2c_one_frame_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_stream -s 25 -c 1 -o gpurun_out/r2c_one_frame -f python tools/one_frame.py 40 > gpurun_out/r2c_one_frame_ncu.log 2>&1
tail -2 gpurun_out/r2c_one_frame_plain.log gpurun_out/r2c_one_frame_ncu.log
timeout 1500 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; tail -c 600 gpurun_out/r2c_bench_ref.json
