#!/bin/bash
# Round-2 evidence run on one B200: tool outputs, the launch list of the bench command and --set full captures of the
# three kernels that changed this round.  Every command runs plain first (exit 0) and only then under ncu.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_tool_outputs.txt
: > $OUT
run() { echo "## $*" >> $OUT; timeout 900 "$@" >> $OUT 2>&1; echo >> $OUT; }
run python tools/small_batch_timing.py 1 8 9 16 32 37 38 64 74 75 128 148
run python tools/cluster_phase_timing.py 1 32
run python tools/config_timings.py
run python tools/config4_timing.py
run bash tools/c4sweep.sh
run python tools/rect_timing.py 1 16
run python tools/mode_timing.py
run python tools/create_timing.py
run python tools/feeder_timing.py 32 120
BENCH="python bench.py --steps 20 --warmup 3 --repeats 3 --no-cpu-baseline --preheat 0 --no-pcie"
$BENCH > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dog_|mode_|flush_|fma_|fadd_" -c 400 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_launches.log 2>&1
$BENCH > gpurun_out/r02_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dog_window45_rot -s 2 -c 1 -f -o gpurun_out/r02_prof_rot $BENCH > gpurun_out/r02_ncu_rot.log 2>&1
python tools/small_batch_timing.py --T 20 32 > gpurun_out/r02_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dog_window45_cluster -s 6 -c 1 -f -o gpurun_out/r02_prof_cluster python tools/small_batch_timing.py --T 20 32 > gpurun_out/r02_ncu_cluster.log 2>&1
python tools/config4_timing.py 64 401 1 > gpurun_out/r02_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dog_rect_argmax_wide -s 3 -c 1 -f -o gpurun_out/r02_prof_wide python tools/config4_timing.py 64 401 1 > gpurun_out/r02_ncu_wide.log 2>&1
ls -la gpurun_out/r02_* | head -30
