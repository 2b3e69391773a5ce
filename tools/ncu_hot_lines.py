#!/usr/bin/env python
"""Warp-stall samples of an .ncu-rep aggregated by CUDA source line (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    data, hdr, fname = [], None, ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 6 or not r[0].strip().isdigit():
            continue
        ki = next(i for i, h in enumerate(hdr) if "Sampl" in h and "Not" not in h and "#" not in h)
        try:
            v = float(r[ki])
        except ValueError:
            continue
        if v > 0:
            # dominant stall reason of the line
            reasons = [(float(r[i]), hdr[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h
                       and r[i].replace(".", "").isdigit()]
            reasons.sort(reverse=True)
            why = ",".join(f"{n[6:]}:{int(c)}" for c, n in reasons[:2] if c > 0)
            data.append((v, fname, r[0], r[1].strip()[:100], why))
    data.sort(reverse=True)
    tot = sum(d[0] for d in data) or 1.0
    print(f"== {rep}: warp-stall samples by source line, {tot:.0f} samples")
    for d in data[:top]:
        print(f"{d[0]:8.0f} {100 * d[0] / tot:5.1f}%  {d[1]}:{d[2]}: {d[3]}   [{d[4]}]")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
