import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f %s  ms/step %.3f  n_gpus %d  scaling %s  launches %s" % (d["value"], d["unit"], d["ms_per_step"], d["n_gpus"], d["scaling"], d.get("gpu_launches")))
print("workload:", d["config"]["workload"][:150], "| frames/step", d["config"]["frames_per_step"], "halos/rank", d["config"].get("halos_per_rank"))
for k in ("roofline", "roofline_correlation"):
    r = d.get(k)
    if r:
        print("%s: %s\n   ms %.4f achieved %.1f %s peak %.1f frac %.3f traffic %s %s" % (k, r["kernel"], r["ms_per_launch"], r["achieved"], r["unit"], r["peak"], r["frac"], r.get("traffic"), r.get("regime", "")))
        if r.get("burst_probe"):
            b = r["burst_probe"]
            print("   burst probe: %d frames ms %.4f achieved %.1f frac %.3f clocks %s" % (b["frames"], b["ms_per_launch"], b["achieved"], b["frac"], b["clocks"].get("sm_mhz")))
for k in ("sustained", "clip_sharding", "frame_sharding"):
    if d.get(k):
        print(k, {a: b for a, b in d[k].items() if a != "clocks"}, (d[k].get("clocks") or {}))
if d.get("roofline_layers"):
    for r in d["roofline_layers"]["rows"]:
        print("  %-52s ms %.4f  %.1f TF/s  frac %.3f  %s" % (r["layer"], r["ms"], r["tflops"], r["frac"], {k: round(v, 4) for k, v in r.items() if k.endswith("_ms")}))
if d.get("roofline_sweep"):
    for r in d["roofline_sweep"]["rows"]:
        print("  %-80s ms %.4f  %s  frac %.3f" % (r["op"], r["ms"], ("%.1f TF/s" % r["tflops"]) if "tflops" in r else ("%.1f GB/s" % r["gbs"]), r["frac"]))
if d.get("e2e"):
    e = d["e2e"]
    print("e2e %.1f frames/s  ms/step %.2f  h2d %.2f GB d2h %.2f GB  ceiling %s frac %.2f  %s" % (e["value"], e["ms_per_step"], e["h2d_bytes_per_step"] / 1e9, e["d2h_bytes_per_step"] / 1e9, e.get("host_copy_ceiling"), e.get("frac_of_host_copy_ceiling", 0), e.get("host_placement")))
print("multi_gpu_check", d.get("multi_gpu_check"))
print("cpu_baseline", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "sample"})
print("clocks", d.get("clocks"))
