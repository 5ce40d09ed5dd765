"""Per-kernel ledger from an `ncu --metrics ... --csv --log-file` capture of one whole pass (profiles/prof_driver.py).

    python profiles/ncu_ledger.py gpurun_out/r2c_metrics_pass.csv [--traffic-json profiles/r2c/conv_dram_traffic.json]

One line per kernel (all its launches aggregated): launches, summed / mean duration, achieved DRAM GB/s and its fraction of the
measured HBM peak (MEASURED_PEAKS.json), L2 throughput, tensor-pipe activity, SM throughput, and a one-word diagnosis.
ncu serialises launches and (by default) flushes caches before each one: durations are cold-cache, compare SHARES."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = dict(t="gpu__time_duration.sum", rd="dram__bytes_read.sum", wr="dram__bytes_write.sum", l2="lts__t_bytes.sum",
         tensor="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", sm="sm__throughput.avg.pct_of_peak_sustained_elapsed",
         dram="gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", warps="sm__warps_active.avg.pct_of_peak_sustained_active",
         issue="sm__inst_issued.avg.pct_of_peak_sustained_active")
UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0, "": 1.0}


def load(path):
    launches = collections.OrderedDict()
    hdr = None
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        v = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
        launches.setdefault(d["ID"], dict(name=name))[d["Metric Name"]] = v
    return list(launches.values())


def main():
    path = sys.argv[1]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else dict(hbm_gbs=6650.0)
    L = load(path)
    agg = collections.OrderedDict()
    for l in L:
        a = agg.setdefault(l["name"], collections.defaultdict(float))
        a["n"] += 1
        for k, m in M.items():
            if m in l:
                a[k] += l[m] * (l.get(M["t"], 0.0) if k in ("tensor", "sm", "dram", "warps", "issue") else 1.0)      # time-weighted percentages
    tot = sum(a["t"] for a in agg.values())
    print(f"# {os.path.basename(path)}: {len(L)} launches, {tot * 1e3:.3f} ms summed kernel time (serialised under ncu; compare shares)")
    print(f"# HBM peak {peaks['hbm_gbs']:.0f} GB/s (MEASURED_PEAKS.json).  Percentages are time-weighted means over a kernel's launches.")
    print("# share%   n  mean_us  DRAM_GB/s (frac)  L2_GB/s  tensor%  sm%  warps%  issue%  diagnosis   kernel")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        t = a["t"]
        gbs = (a["rd"] + a["wr"]) / t / 1e9 if t else 0.0
        l2 = a["l2"] / t / 1e9 if t else 0.0
        pct = {k: (a[k] / t if t else 0.0) for k in ("tensor", "sm", "dram", "warps", "issue")}
        frac = gbs / peaks["hbm_gbs"]
        if pct["tensor"] > 25:
            diag = "tensor"
        elif frac > 0.5:
            diag = "HBM-bound"
        elif pct["issue"] > 55 or pct["sm"] > 55:
            diag = "issue/ALU-bound"
        elif a["t"] / a["n"] < 12e-6:
            diag = "launch/latency (short)"
        else:
            diag = "latency-bound"
        print(f"{100 * t / tot:6.2f} {int(a['n']):4d} {t / a['n'] * 1e6:8.1f}  {gbs:8.0f} ({frac:4.2f})  {l2:7.0f}  {pct['tensor']:6.1f} {pct['sm']:5.1f} {pct['warps']:6.1f} {pct['issue']:6.1f}  {diag:22s} {name}")
    if "--traffic-json" in sys.argv:
        out = sys.argv[sys.argv.index("--traffic-json") + 1]
        conv = [l for l in L if l["name"].startswith("tc_conv")]
        # the UNet evaluation's conv launches = the last 80 (the pass also autotunes nothing under the profiler range)
        d = dict(source=f"{os.path.basename(path)}: ncu --metrics dram__bytes_*.sum,lts__t_bytes.sum over one whole pass (config 2, one DDPM step; cold L2 per launch)",
                 conv_launches=len(conv), dram_read_bytes=sum(l.get(M["rd"], 0.0) for l in conv), dram_write_bytes=sum(l.get(M["wr"], 0.0) for l in conv),
                 l2_bytes=sum(l.get(M["l2"], 0.0) for l in conv), time_us=sum(l.get(M["t"], 0.0) for l in conv) * 1e6,
                 tensor_pipe_active_pct_time_weighted=sum(l.get(M["tensor"], 0.0) * l.get(M["t"], 0.0) for l in conv) / max(1e-30, sum(l.get(M["t"], 0.0) for l in conv)))
        json.dump(d, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
