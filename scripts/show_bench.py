import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
for k in d["kernels"][:int(sys.argv[2]) if len(sys.argv)>2 else 50]: print("   %-22s %.4f %s %.1f"%(k["name"],k["ms"],k.get("bound"),k.get("achieved",0)))
