#!/bin/bash
# final measurement pass of a round: GPU tests + parity report, bench lines, ncu launch list, ncu full capture of the GRU kernel
# usage: bash tools/gpu_final.sh <tag>
T=${1:-r2_final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_tests.txt
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_reference_n1.json 2>> gpurun_out/${T}_bench_n1.err
for cfg in pendulum humanoid; do python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${cfg}_n1.json 2>> gpurun_out/${T}_bench_n1.err; done
python bench.py --config walker2d > gpurun_out/${T}_bench_walker2d_sweep.json 2>> gpurun_out/${T}_bench_n1.err
python bench.py --actor IQN --steps 10 --warmup 3 --no-cpu-baseline --no-ant > gpurun_out/${T}_bench_halfcheetah_iqn_n1.json 2>> gpurun_out/${T}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-ant --no-cpu-baseline --graph 0 > gpurun_out/${T}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_step_v2 -c 2 -o gpurun_out/${T}_gru2_full python tools/gru2_probe.py --reps 1 > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_gru2_full.ncu-rep --page raw --csv > gpurun_out/${T}_gru2_full_raw.csv 2>/dev/null
GAC_GRU_TIMELINE=1 python tools/gru2_probe.py > gpurun_out/${T}_gru2_timeline.txt 2>&1
cat gpurun_out/${T}_tests.txt; head -c 400 gpurun_out/${T}_bench_n1.json; tail -3 gpurun_out/${T}_bench_n1.err
