run() { echo "## $*"; env "$@" timeout 300 python profiles/ab_bundle.py c1 c4 c5s c2 2>&1 | grep '"engine": "auto"' | cut -c1-420; }
run JETS_B200_LIB=$PWD/scratch/libjets_b200_prev.so AB_ROTATE=6
run AB_ROTATE=6
run JETS_B200_LIB=$PWD/scratch/libjets_b200_prev.so AB_ROTATE=6
run AB_ROTATE=6
timeout 300 python -m pytest tests/test_gpu_bundle.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3
