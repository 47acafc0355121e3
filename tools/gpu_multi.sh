#!/bin/bash
# multi-GPU record on one 8-GPU box: bench.py at N = 1, 2, 4, 8 (one process per GPU, NCCL), lines kept under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/multi_gpus.csv 2>&1
cat /sys/fs/cgroup/cpu.max > gpurun_out/multi_cpu_max.txt 2>&1; nproc >> gpurun_out/multi_cpu_max.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/multi_n1.json 2> gpurun_out/multi_n1.err
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/multi_n$N.json 2> gpurun_out/multi_n$N.err
  tail -c 600 gpurun_out/multi_n$N.json; echo
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/multi_n{n}.json").read().strip().splitlines()[-1])
    except Exception as ex:
        print(n, "failed", ex); continue
    if n == 1: base = d
    print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "eff(value)", round(d["value"] / (n * base["value"]), 3) if base else None,
          "eff(e2e)", round(d["e2e"]["value"] / (n * base["e2e"]["value"]), 3) if base else None, "allgather", d.get("allgather_fp16_nchw"))
PY
