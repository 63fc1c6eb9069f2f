"""Summarise scripts/ncu_step_metrics.sh output: per kernel name -> launches, time, share, DRAM GB/s, tensor-pipe %, and a JSON digest."""
import collections, csv, json, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.OrderedDict()   # launch id -> dict
for row in csv.DictReader(lines):
    d = per.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"])})
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
    if m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[m] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tensor_w": 0.0, "xu_w": 0.0, "issue_w": 0.0})
    t = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["us"] += t; a["rd"] += d.get("dram__bytes_read.sum", 0.0); a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    a["tensor_w"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["xu_w"] += t * d.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0.0)
    a["issue_w"] += t * d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
total = sum(a["us"] for a in agg.values())
print("total kernel time %.3f ms over %d launches (serialised, cold-cache ncu replays: compare shares)" % (total / 1e3, len(per)))
print("%9s %6s %5s %9s %9s %9s %8s %7s  %s" % ("ms", "share", "n", "us/launch", "MB/launch", "DRAM GB/s", "tensor%", "issue%", "kernel"))
digest = {}
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    mb = (a["rd"] + a["wr"]) / a["n"] / 1e6
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
    print("%9.3f %5.1f%% %5d %9.1f %9.1f %9.0f %8.1f %7.1f  %s" % (a["us"] / 1e3, 100 * a["us"] / total, a["n"], a["us"] / a["n"], mb, gbs,
                                                               a["tensor_w"] / max(a["us"], 1e-9), a["issue_w"] / max(a["us"], 1e-9), name[:90]))
    digest[name] = {"launches": a["n"], "us_per_launch": a["us"] / a["n"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"],
                    "tensor_pipe_active_pct": a["tensor_w"] / max(a["us"], 1e-9), "share_of_step": a["us"] / total}
if len(sys.argv) > 2:
    g = [v for k, v in digest.items() if "gemm" in k and "tcgen05" in k]
    n = sum(v["launches"] for v in g)
    out = {"source": path, "total_kernel_ms": total / 1e3, "kernels": digest,
           "gemm_family": {"launches": n, "dram_bytes_per_launch": sum(v["dram_bytes_per_launch"] * v["launches"] for v in g) / max(n, 1),
                           "tensor_pipe_active_pct_time_weighted": sum(v["tensor_pipe_active_pct"] * v["us_per_launch"] * v["launches"] for v in g) /
                           max(sum(v["us_per_launch"] * v["launches"] for v in g), 1e-9),
                           "share_of_step": sum(v["share_of_step"] for v in g)}}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print("wrote", sys.argv[2])
