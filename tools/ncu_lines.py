#!/usr/bin/env python
"""Per-source-line summary of an ncu report: joins `ncu --page source --csv` (per SASS
instruction: executed count, stall samples, shared-memory wavefronts) with `nvdisasm -g` line
info of the matching cubin.  Used to write the summaries under profiles/.

    python tools/ncu_lines.py <report.ncu-rep> <object-or-so with the kernel> <kernel substring> [min_pct]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    return [dict(zip(hdr, r)) for r in rows[hdr_i + 1:] if len(r) == len(hdr)]


def mangled_name(rep):
    """mangled name of the (first) kernel in the report, or None"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-kernel-base", "mangled"],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if r]
    try:
        i = rows[0].index("Kernel Name")
        return rows[2][i]
    except (ValueError, IndexError):
        return None


def line_map(binary, kernel_sub, want_len=None):
    """list of (file, line) per instruction index of the function whose name contains kernel_sub; with several
    template instantiations matching, the one whose instruction count equals the profiled one (want_len)"""
    tmp = tempfile.mkdtemp()
    first = None
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
    for cubin in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        cur, fn, res = None, None, {}
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                fn = m.group(1)
                res[fn] = []
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                res[fn].append(cur)
        cands = [(name, lst) for name, lst in res.items() if kernel_sub in name and lst]
        exact = [c for c in cands if want_len is not None and len(c[1]) == want_len]
        if exact or cands:
            found = (exact or cands)[0]
            if first is None or exact:
                first = found
            if exact:
                return first
    if first:
        return first
    raise SystemExit("kernel not found in " + binary)


def main():
    rep, binary, ksub = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
    rows = sass_rows(rep)
    exact = mangled_name(rep)   # the instantiation that was profiled, not the first one that matches the substring
    name, lines = line_map(binary, exact if exact and ksub in exact else ksub, len(rows))
    if len(lines) != len(rows):
        print(f"# warning: {len(rows)} profiled instructions vs {len(lines)} in the binary (different build?)")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r, loc in zip(rows, lines):
        vals = [int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0), int(r["L1 Wavefronts Shared"] or 0)]
        for i in range(3):
            agg[loc][i] += vals[i]
            tot[i] += vals[i]
    print(f"# {name}\n# warp-instructions {tot[0]}  stall samples {tot[1]}  shared wavefronts {tot[2]}")
    src_cache = {}
    for loc in sorted(agg, key=lambda k: (k[0], k[1]) if k else ("", 0)):
        a = agg[loc]
        pct = [100.0 * a[i] / max(tot[i], 1) for i in range(3)]
        if max(pct) < min_pct or loc is None:
            continue
        f, ln = loc
        if f not in src_cache:
            for root, _d, files in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))):
                if f in files:
                    src_cache[f] = open(os.path.join(root, f)).read().splitlines()
                    break
            else:
                src_cache[f] = []
        text = src_cache[f][ln - 1].strip() if ln - 1 < len(src_cache[f]) else ""
        print(f"{f}:{ln:<4d} inst {pct[0]:5.1f}%  smp {pct[1]:5.1f}%  wf {pct[2]:5.1f}%  {text[:90]}")


if __name__ == "__main__":
    main()
