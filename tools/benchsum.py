"""One-screen summary of a bench.py JSON line.   python tools/benchsum.py gpurun_out/r2_bench7.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fwd %.3f ms  value %.3g  dtype %s  e2e %.3g (%.3f ms)  launches %s" % (d["ms_per_step"], d["value"], d["dtype"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("gpu_launches")))
r = d["roofline"]
print("roofline: bound %s frac %.3f | tensor %.1f/%.1f = %.3f (of tf32x3 peak %.3f) | hbm %.0f/%.0f = %.3f" % (r["bound"], r["frac"], r["tensor"]["achieved_tflops"], r["tensor"]["peak_tflops"], r["tensor"]["frac"],
      r["tensor"].get("frac_of_tf32x3_peak") or 0, r["hbm"]["achieved_gbs"], r["hbm"]["peak_gbs"], r["hbm"]["frac"]))
print("clocks:", d.get("clocks"))
for k in ("train", "train_full", "train_lasso"):
    if k in d:
        t = d[k]
        print("%-11s %.3f ms  %.3g samples/s  frac %.3f  kernels %s" % (k, t["ms_per_step"], t["value"], t["roofline"]["frac"],
              {a: round(b, 2) for a, b in t["library_kernel_ms_per_step"].items()}))
if "c1_forward_other_modes" in d:
    print("other modes:", {k: round(v["ms_per_step"], 3) for k, v in d["c1_forward_other_modes"].items() if isinstance(v, dict)})
if "e2e_final_z" in d: print("e2e_final_z %.3f ms" % d["e2e_final_z"]["ms_per_step"])
if "c5" in d:
    c = d["c5"]
    print("c5 all %.2f ms (%.3f of peak %.1f) last_only %.2f ms" % (c["all_iterates"]["ms_per_step"], c["all_iterates"]["frac_of_tensor_peak"], c["tensor_peak_tflops"], c["last_only"]["ms_per_step"]),
          "| 1M:", c.get("total_1m", {}).get("ms"), "| train:", c.get("train", {}).get("ms_per_step"))
if "small_batch" in d: print("small:", [(c["columns"], round(c["fwd_ms"], 3), round(c.get("fwd_ms_back_to_back", 0), 3)) for c in d["small_batch"]["cases"]])
if "safeguard" in d: print("safeguard %.2f ms" % d["safeguard"]["ms"])
if "eager_b200" in d: print("eager:", [(c["variant"], c["columns"], round(c["fwd_ms"], 2), round(c.get("train_ms") or 0, 2)) for c in d["eager_b200"]["cases"]])
print("cpu_baseline:", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
