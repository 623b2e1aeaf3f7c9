set -x
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:jets_fused_fast -s 2 -c 2 -o gpurun_out/r01_c5_fused -f python profiles/prof_fused.py c5 > gpurun_out/prof_c5.log 2>&1
timeout 300 $NCU -k regex:jets_fused_fast -s 2 -c 2 -o gpurun_out/r01_c1_fused -f python profiles/prof_fused.py c1 > gpurun_out/prof_c1.log 2>&1
timeout 300 $NCU -k regex:jets_fused_fast -s 2 -c 2 -o gpurun_out/r01_c2_fused -f python profiles/prof_fused.py c2 > gpurun_out/prof_c2.log 2>&1
timeout 300 $NCU -k regex:gemv -s 2 -c 2 -o gpurun_out/r01_c3_gemv -f python profiles/prof_dense.py > gpurun_out/prof_c3.log 2>&1
ls -la gpurun_out
