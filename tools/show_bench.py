"""Print the parts of a bench JSON line a developer looks at (development aid)."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.4g %s  ms/step %.4f  e2e %.4g  launches %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")))
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k != "step"}, "step", d["roofline"].get("step"))
for k, v in d["kernels"].items():
    print("  %-18s %.4f ms frac %.3f" % (k, v["ms"], v["frac"]))
p = d.get("paths") or {}
for v in p.get("score_scale_variants", []) + [p.get("signal_game_regime_K256")]:
    if v:
        print("K=%d scale %-4g supp %5.1f slots %-5s %.3f ms step_frac %.2f  " % (v["cols"], v["score_scale"], v["support_mean"], v["support_slots"], v["ms_per_step"], v["step_frac"]) +
              " ".join("%s %.3f/%.2f" % (k, x["ms"], x["frac"] or 0) for k, x in v["kernels"].items()))
print("small", {k: {a: round(b, 1) for a, b in v.items() if isinstance(b, float)} for k, v in p.get("small_configs", {}).items()})
for k in ("topk_sparsemax_n128_k10", "topk_sparsemax_marg_n128_k10", "sparsemap_bernoulli_n128", "sparsemap_budget64_n128"):
    if k in p:
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in p[k].items() if not isinstance(b, (dict, list))})
print("sweep", [(b["rows_per_gpu"], round(b["ms_per_step"], 4)) for b in p.get("batch_sweep", [])])
print("allreduce", {a: (round(b, 4) if isinstance(b, float) else b) for a, b in (p.get("sharded_step_with_encoder_allreduce") or {}).items() if not isinstance(b, (dict, list, str))})
print("parity", {k: (v.get("ok") if isinstance(v, dict) else v) for k, v in d["check"]["parity"].items()})
