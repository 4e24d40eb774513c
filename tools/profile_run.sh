#!/bin/bash
# Runs under gpurun: the judged bench (both arms), the serialised ncu launch list and one `--set full` capture per hot
# kernel of the same bench command.  Everything lands in gpurun_out/; tools/profile_extract.py turns it into profiles/.
set -u
mkdir -p gpurun_out
python bench.py --impl reference --steps 200 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
python bench.py --steps 200 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|lpf_|disc_' -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/launches.log 2>&1
for k in pfb256_kernel lpf_phase_kernel disc_recurrence_kernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_$k \
        python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
done
python tools/c3_bench.py 1024 2>/dev/null | grep '^{' > gpurun_out/c3_1024.json
python tools/c3_bench.py 4096 2>/dev/null | grep '^{' > gpurun_out/c3_4096.json
ncu --set full --clock-control none --import-source on -k regex:'symbolize_kernel|sps_fir_kernel' -s 6 -c 2 -f -o gpurun_out/ncu_symbolizer \
    python tools/c3_bench.py 1024 > gpurun_out/ncu_symbolizer.log 2>&1
rm -f gpurun_out/cqpsk_bench.jsonl
for n in 256 1024 4096 8192; do python tools/cqpsk_bench.py $n 2>/dev/null | grep '^{' >> gpurun_out/cqpsk_bench.jsonl; done
python tools/cqpsk_bench.py 1024 48000 10 2>/dev/null | grep '^{' >> gpurun_out/cqpsk_bench.jsonl
python tools/cqpsk_bench.py 1024 48000 8 2>/dev/null | grep '^{' >> gpurun_out/cqpsk_bench.jsonl
ncu --set full --clock-control none --import-source on -k regex:'cqpsk_chain_kernel' -s 4 -c 1 -f -o gpurun_out/ncu_cqpsk_chain_kernel \
    python tools/cqpsk_bench.py 1024 > gpurun_out/ncu_cqpsk_chain_kernel.log 2>&1
python bench.py --workload cqpsk --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_cqpsk.json
python bench.py --workload fec --steps 20 2>/dev/null | tail -1 > gpurun_out/bench_fec.json
ls -la gpurun_out/
