#!/bin/bash
# closing run on N GPUs of one box (N = $1): NCCL test + contract bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist_n$N.log 2>&1; tail -2 gpurun_out/pytest_dist_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/scale_n$N.err | tail -1 > gpurun_out/scale_r1_n$N.json
python3 -c "import json;j=json.load(open('gpurun_out/scale_r1_n$N.json'));print($N,'gpus: %.4g perms/s, %.2f ms/step, e2e %.4g'%(j['value'],j['ms_per_step'],j['e2e']['value']))"
tail -3 gpurun_out/scale_n$N.err
