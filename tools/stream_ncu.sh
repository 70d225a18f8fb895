#!/bin/bash
# ncu capture of the streaming pass kernel:  tools/stream_ncu.sh <tag>
tag=${1:-r02s}
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:k_ppe_stream -s 30 -c 1 -f -o gpurun_out/${tag}_stream $B > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_stream.ncu-rep --page raw --csv > gpurun_out/${tag}_stream_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_stream.ncu-rep --page source --csv > gpurun_out/${tag}_stream_source.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
