import json,sys
d=json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
print("value",round(d["value"],3),"e2e",round(d["e2e"]["value"],3),"launches",d["gpu_launches"],"cpu",d.get("cpu_baseline",{}).get("value"),d.get("cpu_baseline",{}).get("proof_matches_gpu"))
print(d["kernel_ms_per_step"]); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("kernel","achieved","peak","frac","launch_ms","share_of_step")})
print(d["roofline_ntt"]["achieved"], d["roofline_ntt"]["frac"], d["roofline_ntt"]["imad_frac"])
for k,v in d.get("submetrics",{}).items(): print(k,{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
print(d["clocks"])
