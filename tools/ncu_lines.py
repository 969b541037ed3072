#!/usr/bin/env python3
"""Per-source-line aggregation of an ncu SASS source page, using nvdisasm -g line info.
usage: ncu_lines.py <report.ncu-rep> <libptgpu.so> <kernel-substring> [top N] [mangled-symbol-substring]
(the last argument picks one template instantiation in the disassembly, e.g. wf_trace_cw_kernelILb0)"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    out = []
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, loc, active = None, ("?", 0), False
        for line in txt.splitlines():
            m = re.match(r"\s*//-+ \.text\.(\S+)", line)
            if m:
                active = kernel in m.group(1)
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), loc, m.group(2).strip()))
        if out:
            break
    return out


def main():
    rep, so, kernel = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sass = sass_lines(so, sys.argv[5] if len(sys.argv) > 5 else kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the page holds one table per profiled launch; take the first launch of the wanted kernel
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    starts.append(len(rows))
    pick = next(j for j in range(len(starts) - 1) if kernel in rows[starts[j]][1])
    rows = rows[starts[pick]:starts[pick + 1]]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    if len(body) != len(sass):
        print("warning: %d profiled instructions vs %d disassembled" % (len(body), len(sass)))
    ci = {k: hdr.index(k) for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "Source")}
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r, s in zip(body, sass):
        ie, te, sm = int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
        a = agg[s[1]]
        a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
        tot[0] += ie; tot[1] += te; tot[2] += sm
    print("total warp-inst %.3e  thread-inst %.3e  avg active lanes %.2f  samples %d" % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
    # per file summary
    files = collections.defaultdict(lambda: [0, 0, 0])
    for (f, l), a in agg.items():
        files[f][0] += a[0]; files[f][1] += a[1]; files[f][2] += a[2]
    print("-- by file")
    for f, a in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s inst %5.1f%%  lanes %5.2f  samples %5.1f%%" % (f, 100.0 * a[0] / tot[0], a[1] / max(a[0], 1), 100.0 * a[2] / max(tot[2], 1)))
    print("-- top lines by warp instructions")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-20s:%-4d inst %5.2f%%  lanes %5.2f  samples %5.2f%%  (%d sass)" % (f, l, 100.0 * a[0] / tot[0], a[1] / max(a[0], 1), 100.0 * a[2] / max(tot[2], 1), a[3]))


if __name__ == "__main__":
    main()
