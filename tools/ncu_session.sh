#!/bin/bash
# One profiling session on the GPU box (run through gpurun):  tools/ncu_session.sh <tag>
# launch list of the bench step, full captures of the two kernels of a pressure pass and of the non-pressure passes, step case.
tag=${1:-r02}
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary"
$B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ppe_stream -s 30 -c 1 -f -o gpurun_out/${tag}_stream $B > gpurun_out/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:k_ppe_tiled -s 30 -c 1 -f -o gpurun_out/${tag}_frame $B > gpurun_out/${tag}_ncu2b.log 2>&1
ncu --set full --clock-control none -k regex:'k_predict_source|k_correct_rows|k_split_rows' -s 3 -c 3 -f -o gpurun_out/${tag}_other $B > gpurun_out/${tag}_ncu3.log 2>&1
ncu -i gpurun_out/${tag}_stream.ncu-rep --page raw --csv > gpurun_out/${tag}_stream_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_stream.ncu-rep --page source --csv > gpurun_out/${tag}_stream_source.csv 2>/dev/null
ncu -i gpurun_out/${tag}_frame.ncu-rep --page raw --csv > gpurun_out/${tag}_frame_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_other.ncu-rep --page raw --csv > gpurun_out/${tag}_other_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_other.ncu-rep gpurun_out/${tag}_frame.ncu-rep   # gpurun_out/ travels back only below 64 MiB
S="python bench.py --case step --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary"
$S > gpurun_out/${tag}_step_plain.json 2> gpurun_out/${tag}_step_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_step_launches.csv $S > gpurun_out/${tag}_ncu4.log 2>&1
ncu --set full --clock-control none -k regex:k_ppe_tiled -s 30 -c 1 -f -o gpurun_out/${tag}_step_ppe $S > gpurun_out/${tag}_ncu5.log 2>&1
ncu -i gpurun_out/${tag}_step_ppe.ncu-rep --page raw --csv > gpurun_out/${tag}_step_ppe_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_step_ppe.ncu-rep
ls -la gpurun_out | tail -20
