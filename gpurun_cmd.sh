mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -x -q -m gpu) > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
run() { echo "## $*"; env "$@" AB_ROTATE=6 timeout 300 python profiles/ab_bundle.py $CFG 2>&1 | grep '"engine": "auto"' | cut -c1-330; }
CFG="c5s c1 c4 c2"
run X=1
run JETS_B200_STATIC_SCHED=1
