#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an `ncu --set full --import-source on` report.

    python profiles/hot_lines.py REPORT.ncu-rep KERNEL_REGEX [LAUNCH_INDEX] [TOP_N]

ncu's source page (CSV) is per SASS instruction without line numbers; `nvdisasm -gi` of the
cubin inside sigmap_b200/lib/libsigmap_b200.so has the line table.  Both list the kernel's
instructions in the same order, so they are joined by instruction ordinal.
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(rep, kernel, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          f"regex:{kernel}", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1]
    hdr = rows[1]
    ix = {n: k for k, n in enumerate(hdr)}
    body = []
    for r in rows[2:]:
        if len(r) < len(hdr):  # the next kernel's table
            break
        body.append(r)
    return name, ix, body


def line_table(mangled_hint):
    tmp = tempfile.mkdtemp()
    so = os.path.join(ROOT, "sigmap_b200", "lib", "libsigmap_b200.so")
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cubin = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith(".text.") and mangled_hint(l):
            start = i
            break
    table = []
    cur = ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l) or re.match(r"\s+(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+\s", l):
            if "/*" in l or re.match(r"\s+(@!?U?P\d+\s+)?[A-Z]", l):
                table.append(cur)
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    name, ix, rows = sass_rows(rep, kernel, launch)
    base = re.sub(r"\(.*", "", name).split("::")[-1].split("<")[0]
    # template arguments -> their Itanium mangling, e.g. <(bool)0, (bool)1> -> ILb0ELb1EE,
    # <(int)5120, (int)512, (int)4096> -> ILi5120ELi512ELi4096EE: picks the right instantiation
    tmpl = ""
    m = re.search(re.escape(base) + r"<([^>]*)>", name)
    if m:
        parts = []
        for a in m.group(1).split(","):
            k = re.match(r"\s*\((bool|int|unsigned int)\)(-?\d+)", a)
            if not k:
                parts = None
                break
            parts.append({"bool": "Lb", "int": "Li", "unsigned int": "Lj"}[k.group(1)] + k.group(2))
        if parts:
            tmpl = "I" + "E".join(parts) + "EE"
    table = line_table(lambda l: base in l and tmpl in l)
    if len(table) != len(rows):
        print(f"# warning: {len(rows)} SASS rows in the report vs {len(table)} in the cubin "
              f"(rebuilt since the capture?) -- join by ordinal may be off")
    inst, smp = collections.Counter(), collections.Counter()
    stall = collections.defaultdict(collections.Counter)
    stall_cols = [c for c in ix if c.startswith("stall_") and "Not Issued" not in c]
    for k, r in enumerate(rows):
        key = table[k] if k < len(table) else ("?", 0)
        try:
            inst[key] += int(r[ix["Instructions Executed"]])
            smp[key] += int(r[ix["# Samples"]])
        except ValueError:
            continue
        for c in stall_cols:
            try:
                stall[key][c] += int(r[ix[c]])
            except ValueError:
                pass
    ti, ts = sum(inst.values()), sum(smp.values())
    print(f"# {name}\n# {ti:.3e} warp instructions, {ts} stall samples")
    src_cache = {}
    for key, s in smp.most_common(top):
        f, ln = key
        if f not in src_cache:
            p = glob.glob(os.path.join(ROOT, "sigmap_b200", "csrc", f))
            src_cache[f] = open(p[0]).read().splitlines() if p else []
        text = src_cache[f][ln - 1].strip()[:70] if 0 < ln <= len(src_cache[f]) else ""
        top_st = ", ".join(f"{c[6:]} {100 * v / max(s, 1):.0f}%" for c, v in stall[key].most_common(2))
        print(f"{100 * inst[key] / ti:5.1f}% inst {100 * s / ts:5.1f}% smp  {f}:{ln:<4} [{top_st}]  {text}")


if __name__ == "__main__":
    main()
