# round 2: the whole bench line (headline + mixed1m strong scaling + mpc131k + rollout) on N GPUs, launched as the driver does
set -x
N=${1:-8}
TAG=${2:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_${TAG}_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_${N}gpu.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "nccl", d.get("nccl"))
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("unit"), v.get("ms_per_step", v.get("ms_per_tick")), (v.get("e2e") or {}).get("value"), v.get("error"))
PY
ls gpurun_out | head -40
