"""Joins an ncu SASS-level source page (csv) with nvdisasm -g line info: instructions executed per source line.

    cuobjdump -xelf all lib.so; nvdisasm -g -c gk_eval.sm_100a.cubin > eval.sass
    ncu -i rep --page source --csv --kernel-name regex:ac_eval --launch-count 1 > src.csv
    python scripts/sass_by_line.py eval.sass src.csv ac_eval_kernel <units>
"""
import csv, re, sys, collections

sass_path, csv_path, kernel, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
# --- nvdisasm: ordered list of (line) per instruction of the kernel's .text section
lines, cur, on = [], None, False
for l in open(sass_path):
    if l.startswith(".text.") or l.startswith(".section"):
        on = (l.startswith(".text.") and kernel in l)
        continue
    if not on:
        continue
    ms = re.findall(r'File "([^"]+)", line (\d+)', l)
    if ms:
        # with nvdisasm -gi the last pair is the outermost (kernel-level) call site of an inlined function
        cur = (ms[-1][0].split("/")[-1], int(ms[-1][1]))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
rows = list(csv.reader(open(csv_path)))
hdr = next(r for r in rows if "Instructions Executed" in r)
start = rows.index(hdr) + 1
end = next((i for i in range(start, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
rows = rows[:end]
ii, si, pi = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
ti = hdr.index("Thread Instructions Executed")
inst = [(r[si], int(r[ii]), int(r[pi]), int(r[ti])) for r in rows[start:] if len(r) > ii and r[ii].isdigit()]
print("sass instrs: nvdisasm", len(lines), "ncu", len(inst), file=sys.stderr)
agg = collections.OrderedDict()
for k, (src, n, s, t) in enumerate(inst):
    key = lines[k] if k < len(lines) else None
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += n; a[1] += s; a[2] += t
tot = sum(a[0] for a in agg.values())
print("total warp-instructions per unit: %.1f" % (tot / units))
srcs = {}
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if a[0] == 0: continue
    f, ln = key if key else ("?", 0)
    if f not in srcs:
        try: srcs[f] = open("/root/repo/gomokuai_b200/csrc/" + f).read().split("\n")
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%9.1f inst/unit  %5.1f%%  samples %6d  thr/inst %4.1f  %s:%d  %s" % (a[0] / units, 100.0 * a[0] / tot, a[1], a[2] / max(a[0], 1), f, ln, text))
