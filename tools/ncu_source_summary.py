"""Source-level summary of an `ncu --set full --import-source on` report: stall samples and executed instructions
aggregated by CUDA source line, per kernel.

  python tools/ncu_source_summary.py report.ncu-rep regex:v2_pre [top_n] > summary.txt

(`ncu -i report --page source --csv --print-source cuda,sass --kernel-name <filter>` is what it parses; the report itself
is 20 MB per capture and stays out of the repository.)
"""
import collections
import csv
import io
import subprocess
import sys


def summarise(rep: str, kernel_filter: str, top: int = 25) -> str:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel_filter],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = "?", None
    agg = collections.defaultdict(lambda: [0, 0, ""])
    kernel = "?"
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            kernel = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) < 4 or r[2] != "-":   # "-" in the address column = a CUDA source line
            continue
        key = (fname, int(r[0]))
        agg[key][0] += int(r[i_s] or 0)
        agg[key][1] += int(r[i_i] or 0)
        agg[key][2] = r[1].strip()[:100]
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    lines = [f"## {kernel}", f"stall samples {tot_s}, executed warp instructions {tot_i}", "### by stall samples"]
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        lines.append(f"{v[0]:6d} {100 * v[0] / tot_s:5.1f}%  inst {v[1]:8d} {100 * v[1] / tot_i:4.1f}%  {f}:{l}: {v[2]}")
    lines.append("### by executed instructions")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append(f"{v[0]:6d} {100 * v[0] / tot_s:5.1f}%  inst {v[1]:8d} {100 * v[1] / tot_i:4.1f}%  {f}:{l}: {v[2]}")
    return "\n".join(lines)


if __name__ == "__main__":
    print(summarise(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25))
