cp beamforming-lk_b200/libbflk.so /tmp/lib_default.so
for v in default; do
  if [ $v = default ]; then cp /tmp/lib_default.so beamforming-lk_b200/libbflk.so; else cp tools/ubench/libbflk_$v.so beamforming-lk_b200/libbflk.so; fi
  for c in cfg3 cfg2 cfg1 cfg5; do
    python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$v', '$c', round(l['value'],1), round(l['roofline']['frac'],4))"
  done
done
cp /tmp/lib_default.so beamforming-lk_b200/libbflk.so
