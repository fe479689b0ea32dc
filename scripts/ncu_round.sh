#!/bin/bash
# launch list + one ncu --set full capture of a steady step (after the plain run exits 0); the report is
# summarised ON THE BOX (gpurun_out/ only travels back below 64 MiB)
timeout 120 python scripts/fsi_steps.py 4097 10 > gpurun_out/fsi_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
cat gpurun_out/fsi_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 100 --csv --log-file gpurun_out/r02h_launches.csv python scripts/fsi_steps.py 4097 10 > gpurun_out/fsi_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none -s 400 -c 49 -f -o /tmp/r02h_full python scripts/fsi_steps.py 4097 10 > gpurun_out/fsi_ncu2.log 2>&1; echo "full rc=$?"
python scripts/ncu_summary.py /tmp/r02h_full.ncu-rep > gpurun_out/r02h_ncu_full_summary.txt 2> gpurun_out/ncu_summary.err; echo "summary rc=$?"
sz=$(stat -c %s /tmp/r02h_full.ncu-rep); echo "report bytes $sz"
if [ "$sz" -lt 50000000 ]; then cp /tmp/r02h_full.ncu-rep gpurun_out/; fi
