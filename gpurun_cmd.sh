mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:jets_gemm_tc -s 2 -c 2 -o gpurun_out/r01_c3b_tc -f python profiles/prof_dense.py 64 > gpurun_out/prof_c3b.log 2>&1
tail -2 gpurun_out/prof_c3b.log
ls -la gpurun_out/*.ncu-rep
