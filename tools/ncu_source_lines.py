"""Warp-stall samples per CUDA source line of one kernel, from an `ncu --set full --import-source on` report of a -lineinfo build
(the tool behind profiles/r1_response_kernel_source.txt).

Usage: python tools/ncu_source_lines.py report.ncu-rep ["title"] [n_lines]
Runs `ncu -i report --page source --print-source cuda,sass --csv` and aggregates the per-SASS-instruction sampling columns
by source line; prints the overall stall-reason mix and the n_lines hottest lines with their two dominant stall reasons."""
import csv
import subprocess
import sys


def summarize(rep, title="", top=25, out=sys.stdout):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, {}
    for r in csv.reader(txt.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 3 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] != "":              # a source line (its SASS rows follow with an empty first column)
            agg[(cur, int(r[0]))] = (int(r[hdr.index("# Samples")]), r[1].strip(), dict(zip(hdr[4:], r[4:])))
    tot = sum(v[0] for v in agg.values()) or 1
    stalls = {}
    for _, _, d in agg.values():
        for k, x in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and x.isdigit():
                stalls[k[6:]] = stalls.get(k[6:], 0) + int(x)
    out.write(f"# {title or rep}\n# warp-stall samples per CUDA source line; total {tot} samples\n")
    out.write("# stall reasons overall: " + ", ".join(f"{k} {100 * x / tot:.1f}%" for k, x in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]) + "\n")
    for (f, line), (s, text, d) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(((k[6:], int(x)) for k, x in d.items() if k.startswith("stall_") and "Not Issued" not in k and x.isdigit() and int(x) > 0),
                    key=lambda kv: -kv[1])[:2]
        out.write(f"{100 * s / tot:5.1f}%  {f}:{line}  {text[:70]}  [{', '.join(f'{k} {v}' for k, v in st)}]\n")


if __name__ == "__main__":
    summarize(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 25)
