# strong / weak scaling arm of bench.py at N GPUs: usage (under gpurun --gpus N): bash tools/gpu_scale.sh N
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/scale_n${N}_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/scale_n${N}.json 2> gpurun_out/scale_n${N}.err
python - <<PY
import json
d = json.load(open("gpurun_out/scale_n${N}.json"))
print("N", d["n_gpus"], "scaling", d["scaling"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "h2d", d["e2e"]["h2d_bytes_per_step"])
print("other arm", d["config"].get("weak") or d["config"].get("strong"))
print("c4", d["c4"]["value"], d["c4"]["ms"], {k: v["ms"] for k, v in d["c4"]["runs"].items()})
PY
tail -5 gpurun_out/scale_n${N}.err
