"""Per-source-line summary of an ncu report: joins the SASS page of `ncu --page source --csv` (stall samples,
instructions executed) with nvdisasm's line table of the same kernel, in instruction order.

  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> [top N] [section regex, e.g. kernelILi64E]

The cubin is the one inside the shipped .so: `cuobjdump -xelf all mono_vifi_b200/libmonovifi_b200.so`.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    launches, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            launches.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]) - 2:
            cur["rows"].append(r)
    return launches


def line_table(cubin, kernel):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    lines, cur_line, active = [], ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            active = re.search(kernel, ln) is not None
            continue
        if ln.startswith("\t.section"):
            active = False
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = m.group(3)
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)), inl.strip())
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur_line)
    return lines


def main():
    rep, kernel, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    section = sys.argv[5] if len(sys.argv) > 5 else kernel  # regex on the mangled .text section name (template instance)
    launches = sass_rows(rep, kernel)
    if not launches:
        raise SystemExit("no kernel matching %r in %s" % (kernel, rep))
    L = launches[0]
    hdr = L["hdr"]
    si, ii, so = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    table = line_table(cubin, section)
    if len(table) != len(L["rows"]):
        sys.stderr.write("warning: %d SASS rows in the report, %d in the cubin\n" % (len(L["rows"]), len(table)))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    ops = collections.defaultdict(int)
    for r, key in zip(L["rows"], table):
        s, i = int(r[si] or 0), int(r[ii] or 0)
        a = agg[key[:2]]
        a[0] += s
        a[1] += i
        a[2] += 1
        ops[r[so].split()[0] if not r[so].strip().startswith("@") else r[so].split()[1]] += i
    ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print("%s: %d SASS instructions, %d warp-instructions executed, %d stall samples" % (L["name"], len(L["rows"]), ti, ts))
    print("-- by source line (top %d by stall samples)" % top)
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% samples %5.1f%% inst  %4d sass  %s:%d" % (100.0 * a[0] / max(ts, 1), 100.0 * a[1] / max(ti, 1), a[2], key[0], key[1]))
    print("-- by opcode (top 25 by executed warp-instructions)")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
        print("%5.1f%%  %s" % (100.0 * n / max(ti, 1), op))


if __name__ == "__main__":
    main()
