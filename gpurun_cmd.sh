for mx in 1 0; do
  JETS_B200_TC_MIXED=$mx timeout 300 python profiles/ab_tc.py 2>&1 | tail -6
done
for ex in 2 3; do
  AB_SKIP_ACC=1 JETS_B200_TC_EXPT=$ex JETS_B200_TC_MIXED=1 timeout 200 python profiles/ab_tc.py 2>&1 | grep "GB/s"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "dense" 2>&1 | tail -5
