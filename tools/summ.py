import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        if line: print(line[:200])
        continue
    d=json.loads(line); r=d["roofline"]
    print(d["config"]["workload"][:40], "B", d["config"]["frames_per_step"], "maps/s %.1f"%d["value"], r["kernel"], "TF %.2f frac %.3f"%(r["achieved"], r["frac"]), "share %.3f pack %.3f"%(r["kernel_share_of_step"], r["pack_share_of_step"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
