#!/bin/bash
# Runs under gpurun: the judged bench (both arms), the serialised ncu launch list of the same bench command and one
# `--set full` capture per hot kernel.  Everything lands in gpurun_out/; tools/profile_extract.py r02 turns it into profiles/.
# Usage: tools/profile_run_r02.sh [quick]   (quick: only the --set full captures of the kernels named in $KERNELS)
set -u
mkdir -p gpurun_out
KERNELS=${KERNELS:-"lpf_phase_kernel symbolize_kernel sps_fir8_kernel disc_recurrence_kernel p25p1_frame_decode_kernel frame_sync_search_kernel"}
if [ "${1:-full}" != "quick" ]; then
    python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
    python bench.py --steps 200 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench.json
    ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 > gpurun_out/launches.log 2>&1
fi
for k in $KERNELS; do
    ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 7 -c 1 -f -o gpurun_out/ncu_$k \
        python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/ | tail -20
