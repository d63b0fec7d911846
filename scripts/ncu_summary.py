"""Summarise ncu outputs (run here, no GPU): launch list csv -> per-kernel table; .ncu-rep -> key metrics + hot spots."""
import csv, subprocess, sys
from collections import defaultdict

def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    d = defaultdict(list); order = []
    for r in rows[1:]:
        if len(r) > vi:
            d[r[ki]].append(float(r[vi].replace(",", ""))); order.append((r[ki], float(r[vi].replace(",", ""))))
    tot = sum(sum(v) for v in d.values())
    print(f"launch list {path}: total {tot/1e6:.3f} ms")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k[:78]:78s} n={len(v):3d} total={sum(v)/1e6:10.3f} ms ({100*sum(v)/tot:5.1f}%) avg={sum(v)/len(v)/1e3:10.1f} us")
    return order

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "launch__grid_size", "sm__throughput.avg.pct", "dram__cycles_active.avg.pct"]

def report(path, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")][:90])
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS) and not h.endswith(".per_second") and "pct_of_peak_sustained_elapsed" not in h or h in ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"):
                print(f"  {h} [{units[i]}] = {vals[i]}")
        st = {}
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try: st[h.split("stalled_")[1]] = float(vals[i].replace(",", ""))
                except Exception: pass
        tot = sum(st.values()) or 1
        print("  stall samples:", ", ".join(f"{k}={100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    data = []
    for r in rows[2:]:
        try: data.append((int(r[si]), int(r[ii]), r[1].strip()))
        except Exception: pass
    tot = sum(d[0] for d in data) or 1
    print(f"  hot spots (of {tot} samples):")
    for s, i, srcl in sorted(data, reverse=True)[:top]:
        print(f"  {s:8d} {100*s/tot:5.1f}% exec={i:11d} {srcl[:100]}")

if __name__ == "__main__":
    for a in sys.argv[1:]:
        (launches if a.endswith(".csv") else report)(a)
