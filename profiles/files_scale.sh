#!/bin/bash
# BASELINE configs[4] at N GPUs of one box: bench.py --workload files under torchrun, then the
# host ceiling of the same pipeline (PPGS_B200_FILES_NULL_GPU=1).  usage: profiles/files_scale.sh N
N=$1
echo "cores: $(nproc)"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --workload files ${@:2} 2> gpurun_out/r02_files_n${N}.err | tail -1; }
run 29511 > gpurun_out/r02_files_n${N}.json; cut -c1-120 gpurun_out/r02_files_n${N}.json; python -c "
import json; d=json.load(open('gpurun_out/r02_files_n${N}.json')); print('N=${N}', round(d['value']), d['passes_seconds'], d['config']['host_cores'])"
PPGS_B200_FILES_NULL_GPU=1 run 29512 > gpurun_out/r02_files_n${N}_nullgpu.json; python -c "
import json; d=json.load(open('gpurun_out/r02_files_n${N}_nullgpu.json')); print('N=${N} null-gpu', round(d['value']), d['passes_seconds'])"
tail -2 gpurun_out/r02_files_n${N}.err
