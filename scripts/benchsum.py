import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    r=d.get("roofline") or {}
    print("value %.1f e2e %.1f | acc_iso %.3f ms msm_iso %.3f ms frac %.3f inpipe %.3f | launches %s clocks %s cpu %s" % (d["value"], d["e2e"]["value"], r.get("launch_ms_isolated",0), r.get("msm_total_ms_isolated",0), r.get("frac") or 0, r.get("frac_in_pipeline") or 0, d.get("gpu_launches"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value")))
