"""Folds the per-check margin lines written by the GPU parity tests (UHSDR_MARGINS_FILE=... pytest -m gpu) into one JSON:
worst relative error (max|err| / max|ref| on the float audio) and worst SNR per test and case label.
usage: python scripts/collect_margins.py gpurun_out/margins.jsonl profiles/r02_parity_margins.json"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
out = {}
for r in rows:
    k = r["test"].split("::")[-1] if r["test"] else r["label"]
    e = out.setdefault(k, {"label": r["label"], "checks": 0, "max_rel_err": 0.0, "min_snr_db": 1e9})
    e["checks"] += 1
    e["max_rel_err"] = max(e["max_rel_err"], r["max_rel_err"])
    e["min_snr_db"] = min(e["min_snr_db"], r["snr_db"])
summary = {"tolerance": {"max_rel_err": 1e-4, "min_snr_db": 90.0, "reading": "max|err| / max|ref| over the run; SNR = mean(ref^2) / mean(err^2); float audio before int32 formatting"},
           "worst": {"max_rel_err": max(e["max_rel_err"] for e in out.values()), "min_snr_db": min(e["min_snr_db"] for e in out.values()),
                     "at_rel": max(out, key=lambda k: out[k]["max_rel_err"]), "at_snr": min(out, key=lambda k: out[k]["min_snr_db"])},
           "cases": out}
json.dump(summary, open(sys.argv[2], "w"), indent=1, sort_keys=True)
print(json.dumps(summary["worst"]))
