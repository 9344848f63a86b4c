#!/bin/bash
# Round-2 (second part) evidence run on one B200: tool outputs, bench lines, the launch list of the bench command and
# --set full captures of the kernels that changed.  Every command runs plain first (exit 0) and only then under ncu.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02b_tool_outputs.txt
: > $OUT
run() { echo "## $*" >> $OUT; timeout 900 "$@" >> $OUT 2>&1; echo >> $OUT; }
run python tools/tw_sweep.py 8 10 12 15 17 20 22 24 25 26 30 35 50 70 100
run python tools/tw_sweep.py --n 16 10 20 25 30 50 100
run python tools/config4_timing.py
run python tools/config_timings.py
run python tools/small_batch_timing.py 1 32 64 128 148
run python tools/rect_timing.py 1 16
run python tools/issue_probe.py
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_steps20.json 2> gpurun_out/r02b_bench_steps20.err
python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
BENCH="python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu-baseline --preheat 0 --no-pcie"
$BENCH > gpurun_out/r02b_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dog_|mode_|flush_|fma_|fadd_" -c 400 --csv --log-file gpurun_out/r02b_launches.csv $BENCH > gpurun_out/r02b_ncu_launches.log 2>&1
$BENCH > gpurun_out/r02b_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dog_window45_rot -s 2 -c 1 -f -o gpurun_out/r02b_prof_rot $BENCH > gpurun_out/r02b_ncu_rot.log 2>&1
python tools/config4_timing.py 64 401 3 > gpurun_out/r02b_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dog_rows_wide|dog_cols_wide" -s 6 -c 2 -f -o gpurun_out/r02b_prof_twophase python tools/config4_timing.py 64 401 3 > gpurun_out/r02b_ncu_twophase.log 2>&1
python tools/tw_sweep.py 15 > gpurun_out/r02b_plain_tw15.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dog_window45_rot -s 3 -c 1 -f -o gpurun_out/r02b_prof_rot_tw15 python tools/tw_sweep.py 15 > gpurun_out/r02b_ncu_tw15.log 2>&1
ls -la gpurun_out/r02b_* | head -40
