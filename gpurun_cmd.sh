mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_s4.json 2> gpurun_out/bench_n8_s4.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n8_s4.json") if l.startswith("{")][-1])
print("N=8 value", d["value"], "ms/step", d["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["launch_ms"], "e2e", d["e2e"]["value"], "parity", d["parity"])
PY
tail -3 gpurun_out/bench_n8_s4.err
