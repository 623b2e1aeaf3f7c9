# Round-2 profiling recipe (run on the GPU box through gpurun; outputs land in gpurun_out/).
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# (1) the HEADLINE launch itself: bench.py's 48 GB config-5 forward and adjoint (launches 0-4 are warm-up + parity)
timeout 900 $NCU -k regex:jets_fused_bundle -s 5 -c 2 -o gpurun_out/r02_c5_bundle_full -f python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-e2e > gpurun_out/r02_prof_c5_full.log 2>&1
# (2) launch list of the bench command (headline workload only, e2e pipeline included)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_c5.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/r02_bench_under_ncu.log 2>&1
# (3) tcgen05 multi-RHS GEMM, ADJOINT orientation (A' Y): 6 warm-up launches, 10 forward, then the adjoints
timeout 600 $NCU -k regex:jets_gemm_tc -s 17 -c 1 -o gpurun_out/r02_c3b_tc_adjoint -f python profiles/prof_dense.py 64 > gpurun_out/r02_prof_tc_adj.log 2>&1
timeout 600 $NCU -k regex:jets_gemm_tc -s 7 -c 1 -o gpurun_out/r02_c3b_tc_forward -f python profiles/prof_dense.py 64 > gpurun_out/r02_prof_tc_fwd.log 2>&1
# (4) the gated one-launch distributed apply in loopback (block-circulant, one GPU): forward pull, adjoint push
JETS_B200_DIST_LOOPBACK=1 timeout 600 $NCU -k regex:jets_fused_bundle -s 22 -c 2 -o gpurun_out/r02_dist_loopback -f python profiles/prof_dist_loopback.py 32 15625000 2 > gpurun_out/r02_prof_loopback.log 2>&1
ls -la gpurun_out | tail -20
