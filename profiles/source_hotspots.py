"""Per-source-line instruction and stall-sample shares of one kernel launch in an .ncu-rep.

    python profiles/source_hotspots.py REPORT.ncu-rep LAUNCH_INDEX [TOP]

Reads `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo builds and --import-source on).
"""
import csv
import io
import subprocess
import sys


def main():
    rep, k = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr, rows, name = None, None, [], ""
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
        elif r[0] == "Function Name":
            name = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0] != "" and len(r) >= len(hdr) - 2:
            d = dict(zip(hdr, r))  # duplicate "Source" key: the later (SASS) column wins, empty on line rows
            try:
                rows.append((float(d["Instructions Executed"]), float(d["# Samples"]), cur, r[0], r[1].strip()[:88]))
            except ValueError:
                pass
    ti, ts = sum(x[0] for x in rows) or 1, sum(x[1] for x in rows) or 1
    print(name, "| warp instructions", int(ti), "| stall samples", int(ts))
    for x in sorted(rows, reverse=True)[:top]:
        print(f"{100 * x[0] / ti:5.1f}% inst {100 * x[1] / ts:5.1f}% smp  {x[2]}:{x[3]}  {x[4]}")


if __name__ == "__main__":
    main()
