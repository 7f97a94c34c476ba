import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
out=[]
for k,v in (d.get("configs") or {}).items():
    rf=v.get("raw_fastq",{}); so=v.get("soa",{})
    out.append("%s soa %.3f raw %.3f"%(k, so.get("roofline_frac",0), rf.get("roofline_frac",0)))
print(" | ".join(out), "| sm", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
