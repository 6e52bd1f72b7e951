run() { # name, env..., config
  local tag=$1; shift; local cfg=$1; shift
  env "$@" python bench.py --config $cfg --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/s2_$tag.json 2> gpurun_out/s2_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s2_$tag.json")); r = d["roofline"]
    print("$tag", round(d["value"],1), r["kernel"], round(r["frac"],4), "span", d["config"]["tile_span"], "nch", d["config"]["window_chunks"], "pack", round(r["pack_share_of_step"],4))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/s2_$tag.err").read()[-800:])
PY
}
runf() { # tag config frames env...
  local tag=$1; shift; local cfg=$1; shift; local fr=$1; shift
  env "$@" python bench.py --config $cfg --frames $fr --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/s2_$tag.json 2> gpurun_out/s2_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s2_$tag.json")); r = d["roofline"]
    print("$tag", round(d["value"],1), r["kernel"], "frac", round(r["frac"],4), "kernel_share", round(r["kernel_share_of_step"],4), "pack", round(r["pack_share_of_step"],4), "ms", round(d["ms_per_step"],3))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/s2_$tag.err").read()[-800:])
PY
}
