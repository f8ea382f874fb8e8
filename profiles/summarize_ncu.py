"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py full   gpurun_out/prof.ncu-rep      > profiles/<name>_full.txt
  python profiles/summarize_ncu.py launch gpurun_out/launches.csv      > profiles/<name>_launches.txt
  python profiles/summarize_ncu.py traffic gpurun_out/prof.ncu-rep     > profiles/traffic.json   (read by bench.py)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none, one line block per captured launch (%s)" % path)
    for r in data:
        print("\n== launch %s  %s" % (r[idx["ID"]], r[idx["Kernel Name"]]))
        for k in KEYS:
            if k in idx:
                print("   %-66s %16s %s" % (k, r[idx[k]], units[idx[k]]))
        if "dram__bytes_read.sum" in idx:
            def mb(k):
                v, u = float(r[idx[k]]), units[idx[k]]
                return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
            tr = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            t = float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[idx["gpu__time_duration.sum"]]]
            print("   %-66s %16.1f MB  -> %.0f GB/s over the launch" % ("traffic = dram read + write", tr, tr / t * 1e3))


def launch(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void dcrf::<unnamed>::", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES) %s" % path)
    print("%-60s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.1f %10.2f %6.1f%%" % (name[:60], n, t, t / n, 100 * t / tot))


def traffic(path):
    """DRAM bytes per launch per kernel class of the 1080p bench workload, stamped with the hash of the kernel sources the
    capture was taken from (bench.py refuses the file when that hash differs from the build it is timing)."""
    import json
    import os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cls = collections.OrderedDict()
    for r in data:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void dcrf::<unnamed>::", "")
        c = ("splat" if "k_splat" in name else "blur2" if "k_blur2" in name else "blur" if "k_blur<" in name
             else "fused_update" if "k_fused_update" in name else None)
        if c is None or "<(int)1" in name or "<(int)0, (int)0" in name:   # skip the norm pass and the Q0 softmax
            continue
        b = sum(float(r[idx[k]]) * scale[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        e = cls.setdefault(c, {"kernel": short, "bytes": [], "us": []})
        e["bytes"].append(b)
        e["us"].append(float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[idx["gpu__time_duration.sum"]]])
    out = {"_kernel_sources_sha16": bench.kernel_sources_sha(),
           "_source": "%s: ncu --set full --clock-control none over the 1080p bench workload; dram__bytes_read.sum + "
                      "dram__bytes_write.sum per launch, averaged over the captured launches of the kernel class" % os.path.basename(path)}
    for c, e in cls.items():
        out[c] = {"kernel": e["kernel"], "dram_bytes_per_launch": sum(e["bytes"]) / len(e["bytes"]),
                  "launches_captured": len(e["bytes"]), "avg_us_under_ncu": sum(e["us"]) / len(e["us"])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"full": full, "launch": launch, "traffic": traffic}[sys.argv[1]](sys.argv[2])
