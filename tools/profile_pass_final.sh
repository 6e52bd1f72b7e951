#!/bin/bash
# End-of-round profiling pass (gpurun -- bash tools/profile_pass_final.sh): the default bench first, outside any profiler; then the
# ncu launch list of the same command, one `--set full` capture of the dominant kernel at the default command, and one of the
# channel-split (thread-block cluster) kernel a single live cfg3 frame runs.  Summaries land in gpurun_out/.
set -u
mkdir -p gpurun_out
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cfg3.csv python bench.py $B > gpurun_out/final_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:das_tile_kernel -s 2 -c 1 -o gpurun_out/final_full_cfg3 -f python bench.py $B > gpurun_out/final_ncu_cfg3.log 2>&1
python tools/ncu_summary.py gpurun_out/final_full_cfg3.ncu-rep --stalls > gpurun_out/final_ncu_fma2_cfg3.txt 2>&1
BFLK_LAT_WARPS=8 BFLK_LAT_SPLIT=4 ncu --set full --clock-control none --import-source on -k regex:das_tile_kernel -s 20 -c 1 -o gpurun_out/final_full_split -f python tools/latency_sweep.py cfg3 8x4 > gpurun_out/final_ncu_split.log 2>&1
python tools/ncu_summary.py gpurun_out/final_full_split.ncu-rep --stalls > gpurun_out/final_ncu_split_cfg3.txt 2>&1
rm -f gpurun_out/final_full_split.ncu-rep
cat gpurun_out/final_ncu_fma2_cfg3.txt gpurun_out/final_ncu_split_cfg3.txt
