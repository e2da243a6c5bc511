#!/bin/bash
# round-2 measurement pass: v2 GRU timeline + probes, bench line, ncu launch list, ncu full capture of the v2 kernel
mkdir -p gpurun_out
{
GAC_GRU_TIMELINE=1 python tools/gru2_probe.py
for p in 1 2 4 8 16 32 64 31; do echo "--- probe $p"; GAC_G2_PROBE=$p GAC_GRU_TIMELINE=1 python tools/gru2_probe.py 2>&1 | tail -4; done
} > gpurun_out/r2b_gru2_timeline.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-ant --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --no-ant --no-cpu-baseline --graph 0 > gpurun_out/r2b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_step_v2 -c 2 -o gpurun_out/r2b_gru2_full python tools/gru2_probe.py --reps 1 > gpurun_out/r2b_ncu2.log 2>&1
ncu -i gpurun_out/r2b_gru2_full.ncu-rep --page raw --csv > gpurun_out/r2b_gru2_full_raw.csv 2>/dev/null
tail -3 gpurun_out/r2b_gru2_timeline.txt
