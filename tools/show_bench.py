import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["gpu_launches"]); r=d["roofline"]
for k,v in r["classes"].items(): print("  ",k,round(v["ms"],2),round(v["alg_GBs"]))
