#!/bin/bash
# host bandwidth probe at 1/2/4/8 ranks, then the bench at 8 GPUs with the round driver's flags
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) \
      scripts/host_bandwidth.py 2>/dev/null | grep aggregate >> gpurun_out/host_bw.jsonl
done
cat gpurun_out/host_bw.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_k20.json 2> gpurun_out/bench_8gpu_k20.err
tail -2 gpurun_out/bench_8gpu_k20.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_8gpu_k20.json"))
print("value %.3e e2e %.3e"%(d["value"], d["e2e"]["value"]), d["clocks"])
PY
