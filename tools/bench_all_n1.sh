# round-1 refresh of every single-GPU configuration (BASELINE.json configs); raw JSON lines -> gpurun_out/r1d_*.json
set -x
python bench.py > gpurun_out/r1d_dragon_orbit_n1.json 2> gpurun_out/r1d_dragon_orbit_n1.err
python bench.py --impl reference > gpurun_out/r1d_reference_arm.json 2> gpurun_out/r1d_reference_arm.err
for w in dragon_closeup_960x540 bunny_960x540 happy_orbit_3840x2160 synthetic10m_7680x4320; do
  python bench.py --workload $w --steps 5 > gpurun_out/r1d_${w}_n1.json 2> gpurun_out/r1d_${w}_n1.err
done
tail -c 300 gpurun_out/r1d_*_n1.err
