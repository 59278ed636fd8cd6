import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fwd", round(d["value"]/1e6,2), "M/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]/1e6,2), {k:round(v,3) for k,v in d["roofline"]["per_kind_ms_per_step"].items()})
t=d["train"]; print("train", round(t["value"]/1e6,2), "M/s", round(t["ms_per_step"],2), "ms", {k:round(v,2) for k,v in t["library_kernel_ms_per_step"].items()})
