"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS view) with nvdisasm line info.
usage: python scripts/ncu_lines.py <rep.ncu-rep> <object.o> <mangled kernel name> [top]"""
import csv, re, subprocess, sys, os, tempfile, collections
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp(dir=os.path.dirname(os.path.abspath(rep)))
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(kern + ":"))
lines = []  # (offset, file:line, text)
cur = "?"
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the page holds one section per captured launch: take the first one (pass -k / -c to ncu to choose the kernel)
sec = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = sec[int(os.environ.get("NCU_SEC", "0")):]  # NCU_SEC=k: the k-th captured launch
end = sec[1] if len(sec) > 1 else len(rows)
print("kernel:", rows[sec[0]][1])
hdr = rows[sec[0] + 1]; ix = {h: i for i, h in enumerate(hdr)}
body = rows[sec[0] + 2:end]
base = int(body[0][0], 16)
off2line = {o: (ln, t) for o, ln, t in lines}
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0; totinst = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in body:
    off = int(r[0], 16) - base
    ln, _ = off2line.get(off, ("?", ""))
    s = int(r[ix["# Samples"]] or 0); ie = int(r[ix["Instructions Executed"]] or 0)
    agg[ln][0] += s; agg[ln][1] += ie; tot += s; totinst += ie
    for h in stall_cols:
        v = int(r[ix[h]] or 0)
        if v: agg[ln][2][h] += v
print("total samples", tot, "warp instructions", totinst)
for ln, (s, ie, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%6.2f%% samples %6.2f%% inst  %-28s %s" % (100.0 * s / tot, 100.0 * ie / totinst, ln, ", ".join("%s %d" % (k[6:], v) for k, v in st.most_common(3))))
import shutil; shutil.rmtree(tmp)
