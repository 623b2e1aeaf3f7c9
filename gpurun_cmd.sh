set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
PROF_ENGINE=auto timeout 300 $NCU -k regex:jets_fused -s 2 -c 1 -o gpurun_out/c4_bundle -f python profiles/prof_fused.py c4 > gpurun_out/prof_c4b.log 2>&1
PROF_ENGINE=tma_nocache timeout 300 $NCU -k regex:jets_fused -s 2 -c 1 -o gpurun_out/c4_old -f python profiles/prof_fused.py c4 > gpurun_out/prof_c4o.log 2>&1
tail -3 gpurun_out/prof_c4b.log gpurun_out/prof_c4o.log
ls -la gpurun_out
