# The three multi-GPU BASELINE configurations on 8 GPUs of one box (one process per GPU under torchrun): usage  bash tools/bench_n8.sh [tag]
TAG=${1:-r02s}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$TR 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n8_large.json 2> gpurun_out/${TAG}_n8_large.err; tail -c 400 gpurun_out/${TAG}_n8_large.err
$TR 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --model small --quant int8 --chunks 16 > gpurun_out/${TAG}_bench_n8_small_int8.json 2> gpurun_out/${TAG}_n8_small.err; tail -c 400 gpurun_out/${TAG}_n8_small.err
$TR 29513 bench.py --gpus 8 --steps 5 --warmup 3 --workload streaming --model medium --quant int4 --chunks 64 --streams 64 > gpurun_out/${TAG}_bench_n8_streaming.json 2> gpurun_out/${TAG}_n8_stream.err; tail -c 400 gpurun_out/${TAG}_n8_stream.err
python - <<PY
import json
for f in ["large","small_int8","streaming"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/$TAG" + f"_bench_n8_{f}.json") if l.startswith("{")][-1])
        print(f, d["n_gpus"], round(d["value"]), (d.get("e2e") or {}).get("value"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
