"""Per-source-line profile of one kernel launch from an ncu report captured with --import-source on (binary built with -lineinfo):
ncu's CLI source page is per SASS instruction; nvdisasm -g on the SAME cubin gives every instruction's file / line (and inline chain), so
the two are joined by instruction index. Lines are attributed to the outermost call site in the kernel body (innermost with --inner).

    cuobjdump -xelf all par_raytracer_b200/librt_b200.so; nvdisasm -g -c rt_render.sm_100a.cubin > render.sass
    ncu -i X.ncu-rep --page source --csv > src.csv
    python scripts/ncu_by_line.py render.sass src.csv <mangled kernel prefix> <kernel index in the report> [--inner] [--top N]
"""
import collections, csv, re, sys

sass_file, csv_file, mangled, ki = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
inner = "--inner" in sys.argv
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
txt = open(sass_file).read().split("\n")
start = [n for n, l in enumerate(txt) if l.startswith(".text." + mangled)][0]
seq = []; cur = ("?", 0, "")
for l in txt[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and seq:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(csv_file)))
kern = []; c = None
for r in rows:
    if r and r[0] == "Kernel Name":
        c = {"name": r[1], "rows": []}; kern.append(c)
    elif r and r[0] == "Address":
        c["hdr"] = r
    elif c is not None and r:
        c["rows"].append(r)
k = kern[ki]; H = k["hdr"]
assert len(k["rows"]) == len(seq), (len(k["rows"]), len(seq), "report and cubin are different builds")
ii, it, iss = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
by = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()]); tot = ts = 0
allst = collections.Counter()
for r, (ins, (f, ln, extra)) in zip(k["rows"], seq):
    key = (f, ln)
    m = re.findall(r'inlined at "([^"]+)", line (\d+)', extra)
    if m and not inner:
        key = (m[-1][0].split("/")[-1], int(m[-1][1]))
    cnt = int(r[ii]); b = by[key]; b[0] += cnt; b[1] += int(r[it]); b[2] += int(r[iss]); tot += cnt; ts += int(r[iss])
    for s in stalls:
        v = int(r[H.index(s)] or 0); b[3][s] += v; allst[s] += v
print(f"# {k['name'][:60]}: {tot} warp instructions, {sum(b[1] for b in by.values()) / tot:.1f} active lanes per instruction, {ts} stall samples")
print("# stall reasons (% of samples): " + ", ".join(f"{s.replace('stall_', '')} {100 * v / ts:.1f}" for s, v in allst.most_common(8)))
srcs = {}
for (f, ln), (cnt, th, sm, st) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open("par_raytracer_b200/csrc/" + f).read().split("\n")
        except Exception:
            srcs[f] = []
    line = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    top3 = " ".join(f"{s.replace('stall_', '')}:{100 * v // max(1, sum(st.values()))}" for s, v in st.most_common(2))
    print(f"{f:14s} L{ln:<4d} inst {100 * cnt / tot:5.1f}%  lanes {th / max(cnt, 1):4.1f}  samples {100 * sm / ts:5.1f}%  {top3:28s} | {line}")
