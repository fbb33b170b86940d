#!/usr/bin/env python
"""profiles/r2_traffic.json from the ncu captures of profiles/tools/traffic_run.py (r2_capture.sh):
DRAM bytes and kernel time per traversal of the CLV launches (k_clv*, k_cherry*) and, for the site-repeat
configuration, of the identifier launches (k_rid*, k_rep*).

  python profiles/tools/traffic_json.py <dir with r2_traffic_<config>.csv and tr_<config>.log> > profiles/r2_traffic.json"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from traffic_summary import load  # noqa: E402

KEYS = {"dna": "dna_100x{sites}", "aa": "aa_lg4m_200x{sites}", "repeats": "repeats_1000x{sites}",
        "repeats_ids": "repeats_ids_1000x{sites}"}


def main(d):
    out = {"how": "ncu --kernel-name regex:'k_clv|k_cherry|k_rid|k_rep' --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                  "dram__bytes_write.sum --clock-control none over profiles/tools/traffic_run.py <config> (PLF_GRAPH=0): one "
                  "set-up traversal + 3 more; DRAM bytes of the CLV launches divided by the traversals captured"}
    for cfg, key in KEYS.items():
        csv_path, log_path = os.path.join(d, f"r2_traffic_{cfg}.csv"), os.path.join(d, f"tr_{cfg}.log")
        if not os.path.exists(log_path):
            log_path = os.path.join(d, f"r2_traffic_{cfg}.log")  # as committed under profiles/
        if not (os.path.exists(csv_path) and os.path.exists(log_path)):
            continue
        info = None
        for line in open(log_path):
            if line.startswith("{"):
                info = json.loads(line)
        if info is None:
            continue
        rows = load(csv_path)
        clv = {"n": 0, "us": 0.0, "bytes": 0.0}
        ids = {"n": 0, "us": 0.0, "bytes": 0.0}
        per_kernel = {}
        for (_, name), m in rows.items():
            tgt = ids if re.match(r"k_rid|k_rep", name) else clv
            b = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
            tgt["n"] += 1
            tgt["us"] += m.get("gpu__time_duration.sum", 0)
            tgt["bytes"] += b
            k = per_kernel.setdefault(name[:70], [0, 0.0, 0.0])
            k[0] += 1
            k[1] += m.get("gpu__time_duration.sum", 0)
            k[2] += b
        trav = info["traversals"] + 1  # the set-up traversal is captured too
        id_trav = trav if cfg == "repeats_ids" else 1
        rec = {"traffic_bytes_per_step": round(clv["bytes"] / trav), "clv_kernel_us_per_step_under_ncu": round(clv["us"] / trav, 1),
               "clv_launches_per_step": clv["n"] / trav, "traversals_captured": trav,
               "source": f"profiles/r2_traffic_{cfg}.csv", "logl": info["logl"],
               "per_kernel": {k: {"launches": v[0], "us": round(v[1], 1), "dram_gb": round(v[2] / 1e9, 3),
                                  "dram_gbs": round(v[2] / v[1] / 1e3, 1) if v[1] else None}
                              for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])}}
        if ids["n"]:
            rec["identifier_bytes_per_update"] = round(ids["bytes"] / id_trav)
            rec["identifier_kernel_us_per_update_under_ncu"] = round(ids["us"] / id_trav, 1)
        out[key.format(sites=info["sites"])] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
