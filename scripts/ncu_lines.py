"""Per-source-line stall samples from an .ncu-rep captured with --import-source on (run here, no GPU).
Usage: python scripts/ncu_lines.py file.ncu-rep [min_pct]"""
import csv, subprocess, sys

def main(path, min_pct=0.4):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur_file, hdr, data = None, None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].strip().isdigit():
            try:
                data.append((int(r[hdr.index("# Samples")]), cur_file, int(r[0]), r[1].strip()))
            except Exception:
                pass
    tot = sum(d[0] for d in data) or 1
    print(f"{path}: {tot} samples")
    for s, f, ln, src in sorted(data, key=lambda d: (d[1], d[2])):
        if 100.0 * s / tot >= min_pct:
            print(f"{s:8d} {100*s/tot:5.1f}%  {f}:{ln:<4d} {src[:120]}")

if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4)
