# Round-1 profiling recipe (run on the GPU box through gpurun; outputs land in gpurun_out/).
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:jets_fused_bundle -s 2 -c 2 -o gpurun_out/r01_c5_bundle -f python profiles/prof_fused.py c5 > gpurun_out/prof_c5.log 2>&1
timeout 300 $NCU -k regex:jets_fused_bundle -s 2 -c 2 -o gpurun_out/r01_c1_bundle -f python profiles/prof_fused.py c1 > gpurun_out/prof_c1.log 2>&1
timeout 300 $NCU -k regex:jets_fused_bundle -s 2 -c 2 -o gpurun_out/r01_c2_bundle -f python profiles/prof_fused.py c2 > gpurun_out/prof_c2.log 2>&1
# launch list of the bench command itself (headline workload only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_c5.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out
