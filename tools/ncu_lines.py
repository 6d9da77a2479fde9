"""Attribute ncu warp-stall samples (SASS view of a .ncu-rep) to source lines.

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [CUBIN] [--top N] [--id K]

The SASS page of a report captured without --import-source carries per-instruction
samples but no line numbers; `nvdisasm -g` on the cubin of the same build gives the
line of every instruction in the same order, so the two are joined by position.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def sass_samples(rep, regex, kid=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex]
    if kid is not None:
        cmd += ["--launch-skip", str(kid), "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel only
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    return rows[hdr_i - 1][1] if hdr_i else "", hdr, body


def cubin_lines(cubin, func_regex):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines = []
    cur_fn, active, cur_line = None, False, ("?", 0)
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            active = re.search(func_regex, cur_fn) is not None
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append((cur_line, ln.strip()))
    return lines


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    cubin = "/tmp/cub/em.sm_100a.cubin"
    top, kid = 25, None
    args = sys.argv[3:]
    while args:
        a = args.pop(0)
        if a == "--top":
            top = int(args.pop(0))
        elif a == "--id":
            kid = int(args.pop(0))
        else:
            cubin = a
    name, hdr, body = sass_samples(rep, regex, kid)
    ci = hdr.index("Warp Stall Sampling (All Samples)")
    ii = hdr.index("Instructions Executed")
    lines = cubin_lines(cubin, regex)
    print(name, "SASS rows", len(body), "cubin instrs", len(lines))
    n = min(len(body), len(lines))
    agg = defaultdict(lambda: [0, 0])
    tot = 0
    for k in range(n):
        s = int(body[k][ci] or 0)
        agg[lines[k][0]][0] += s
        agg[lines[k][0]][1] += int(body[k][ii] or 0)
        tot += s
    print("total samples", tot)
    for (f, l), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * s / max(tot, 1):6.2f}%  {s:7d} samples {i:10d} instr  {f}:{l}")


if __name__ == "__main__":
    main()
