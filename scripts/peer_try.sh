#!/bin/bash
# 2-rank checks of the peer-memory exchanges: both ranks on one GPU (gloo set-up), then one rank per GPU
NP=${NP:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
if [ "$NP" = "2" ]; then
RMT_SAME_GPU=1 timeout 300 $TR --master-port 29521 scripts/slab_check.py --check 513 --fsi 513 --pfsi 513 > gpurun_out/peer_same.json 2> gpurun_out/peer_same.err
echo "same-gpu rc=$?"; cat gpurun_out/peer_same.json; tail -5 gpurun_out/peer_same.err | grep -v "^\*\|OMP_NUM\|^$"
RMT_SLAB_COMM=nccl timeout 300 $TR --master-port 29524 scripts/slab_check.py --check 0 --pfsi 1025 > gpurun_out/nccl_two.json 2> gpurun_out/nccl_two.err
echo "nccl rc=$?"; cat gpurun_out/nccl_two.json
fi
timeout 300 $TR --master-port 29522 scripts/slab_check.py --check 1025 --fsi 1025 --pfsi 1025 --fsi-time 4097 --pfsi-time 8193 --pfluid-time 16385 > gpurun_out/peer_$NP.json 2> gpurun_out/peer_$NP.err
echo "peer rc=$?"; cat gpurun_out/peer_$NP.json; tail -5 gpurun_out/peer_$NP.err | grep -v "^\*\|OMP_NUM\|^$"
timeout 300 $TR --master-port 29523 scripts/slab_profile.py 4097 > gpurun_out/slabprof_$NP.txt 2>gpurun_out/slabprof_$NP.err
echo "profile rc=$?"; head -3 gpurun_out/slabprof_$NP.txt; rm -f gpurun_out/slab_trace_r0.json
