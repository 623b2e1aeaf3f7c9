run() { echo "## $*"; env "$@" timeout 300 python profiles/ab_bundle.py c5 2>&1 | grep '"engine": "auto"' | cut -c1-330; }
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
run X=1
run JETS_B200_STATIC_SCHED=1
run JETS_B200_NO_PDL=1
run X=2
