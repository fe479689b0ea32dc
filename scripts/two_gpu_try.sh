#!/bin/bash
export RMT_PEER_TIMEOUT=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29551 bench.py --gpus 2 --no-slab --steps 10 > gpurun_out/b2q.json 2> gpurun_out/b2q.err; echo "bench rc=$?"
grep "^\[rank0\]" gpurun_out/b2q.err | grep -v "\^\^" | tail -6
RMT_PEER_DISABLE_IPC=1 timeout 200 $TR --master-port 29552 scripts/slab_check.py --check 513 --fsi 513 > gpurun_out/fallback.json 2> gpurun_out/fallback.err; echo "fallback rc=$?"; cat gpurun_out/fallback.json; grep "unavailable" gpurun_out/fallback.err
