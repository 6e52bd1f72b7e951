#!/bin/bash
# One profiling pass on the GPU box (gpurun -- bash tools/profile_pass.sh): the bench first, outside any profiler; then the
# ncu launch list of the same command and one `--set full` capture of the dominant kernel per configuration.  Everything
# lands in gpurun_out/; the summaries are copied to profiles/ by hand (tools/ncu_summary.py).
set -u
mkdir -p gpurun_out
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
python bench.py --steps 20 --warmup 5 > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches_cfg3.csv python bench.py $B > gpurun_out/prof_ncu_list.log 2>&1
for cfg in cfg3 cfg1; do
  ncu --set full --clock-control none --import-source on -k regex:das_tile_kernel -s 2 -c 1 -o gpurun_out/prof_full_$cfg -f python bench.py --config $cfg $B > gpurun_out/prof_ncu_$cfg.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_full_$cfg.ncu-rep --stalls > gpurun_out/prof_ncu_fma2_$cfg.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:das_tile_kernel -s 2 -c 1 -o gpurun_out/prof_full_cfg3_exact -f python bench.py --config cfg3 --kernel 2 $B > gpurun_out/prof_ncu_cfg3_exact.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_full_cfg3_exact.ncu-rep --stalls > gpurun_out/prof_ncu_exact_cfg3.txt 2>&1
ncu --set full --clock-control none -k regex:pack_kernel -s 1 -c 1 -o gpurun_out/prof_full_pack -f python bench.py $B > gpurun_out/prof_ncu_pack.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_full_pack.ncu-rep > gpurun_out/prof_ncu_pack_cfg3.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:miso_kernel -s 40 -c 1 -o gpurun_out/prof_full_miso -f python tools/miso_time.py > gpurun_out/prof_ncu_miso.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_full_miso.ncu-rep --stalls > gpurun_out/prof_ncu_miso_cfg4.txt 2>&1
rm -f gpurun_out/prof_full_cfg1.ncu-rep gpurun_out/prof_full_pack.ncu-rep gpurun_out/prof_full_miso.ncu-rep gpurun_out/prof_full_cfg3_exact.ncu-rep   # keep one report (size)
cat gpurun_out/prof_ncu_fma2_cfg3.txt gpurun_out/prof_ncu_fma2_cfg1.txt gpurun_out/prof_ncu_exact_cfg3.txt gpurun_out/prof_ncu_pack_cfg3.txt gpurun_out/prof_ncu_miso_cfg4.txt
