#!/bin/bash
# usage: tools/ab_bench.sh <out file> <name=path/to/libbflk.so> ...  -- same-box comparison of library builds: bench.py
# (kernel-only leg) for cfg3 / cfg2 / cfg1 / cfg5 with each build in turn; the in-tree library is restored afterwards.
out=$1; shift
lib=beamforming-lk_b200/libbflk.so
cp $lib /tmp/libbflk_keep.so
for spec in "$@"; do
  name=${spec%%=*}; path=${spec#*=}
  [ "$path" != "$lib" ] && cp "$path" $lib
  for c in ${AB_CONFIGS:-cfg3 cfg2 cfg1 cfg5}; do
    python bench.py --config $c --steps ${AB_STEPS:-20} --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null |
      python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '$c', round(l['value'],1), 'frac', round(l['roofline']['frac'],4), 'ms', round(l['ms_per_step'],3))" >> $out
  done
  cp /tmp/libbflk_keep.so $lib
done
cat $out
