mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_bundle.py -x -q -m gpu) > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
run() { echo "## $*"; env "$@" AB_ROTATE=6 timeout 300 python profiles/ab_bundle.py $CFG 2>&1 | grep '"engine": "auto"' | cut -c1-330; }
CFG="c5s c1 c4 c2"
run X=1
run JETS_B200_FAST_VARIANT=3
run JETS_B200_FAST_VARIANT=4
run JETS_B200_FAST_VARIANT=5
