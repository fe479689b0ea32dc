#!/bin/bash
# one GPU call: the GPU test suite, then the default bench line (1 GPU)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gputests.log
timeout 600 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench1.err
timeout 100 python scripts/graph_probe.py 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
