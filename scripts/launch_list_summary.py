"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv): per kernel launches, total time, share, DRAM bytes per launch.
usage: python scripts/launch_list_summary.py launches.csv "header" summary.txt traffic.json workload"""
import collections
import csv
import json
import re
import sys

path, header, out_txt, out_json, workload = sys.argv[1:6]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iID], dict(name=r[iK]))
    d[r[iM]] = float(r[iV].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    name = re.sub(r"\(.*", "", d["name"]).replace("at::", "").strip()[:70]
    a = agg.setdefault(name, dict(n=0, ns=0.0, rd=0.0, wr=0.0))
    a["n"] += 1
    a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a["ns"] for a in agg.values())
with open(out_txt, "w") as f:
    f.write(header + "\n")
    f.write("%-72s %8s %12s %8s %14s %14s\n" % ("kernel", "launches", "total ms", "share", "DRAM read/launch", "DRAM write/launch"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])[:30]:
        f.write("%-72s %8d %12.3f %7.2f%% %11.1f MB %11.1f MB\n" % (
            name, a["n"], a["ns"] / 1e6, 100 * a["ns"] / tot, a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6))
    f.write("total %.1f ms over %d launches\n" % (tot / 1e6, len(per)))
ops = {"tc_fringe_fwd_kernel": "tcfringe_fwd", "tc_fringe_bwd_kernel": "tcfringe_bwd",
       "ant_fringe_fwd_kernel": "antfringe_fwd", "ant_fringe_bwd_kernel": "antfringe_bwd",
       "build_interp_t_kernel": "build_interp_t", "build_interp_bwd_kernel": "build_interp_bwd",
       "reduce_units_kernel": "reduce_units", "interp_transpose_kernel": "interp_transpose",
       "fringe_sum_fwd_kernel": "fringe_sum_fwd", "fringe_sum_bwd_sky_kernel": "fringe_sum_bwd_sky"}
kern = {}
for name, a in agg.items():
    for pat, op in ops.items():
        if pat in name:
            kern[op] = dict(kernel=name, launches=a["n"], ms_per_launch=a["ns"] / a["n"] / 1e6,
                            dram_bytes_per_launch=(a["rd"] + a["wr"]) / a["n"],
                            dram_read_per_launch=a["rd"] / a["n"], dram_write_per_launch=a["wr"] / a["n"])
json.dump(dict(workload=workload, note=header, kernels=kern), open(out_json, "w"), indent=1)
