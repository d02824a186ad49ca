#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/gpus.txt
timeout 200 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist8.log 2>&1; tail -2 gpurun_out/pytest_dist8.log
for n in 8 4 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 2 2> gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_r1_n$n.json
  python3 -c "import json;j=json.load(open('gpurun_out/scale_r1_n$n.json'));print($n,'gpus: %.4g perms/s, %.2f ms/step, e2e %.4g'%(j['value'],j['ms_per_step'],j['e2e']['value']))"
done
