"""Aggregate the `--page source --print-source cuda,sass` CSV of an .ncu-rep per CUDA source line:
stall samples, executed instructions, top stall reasons.  python tools/ncu_lines.py rep.ncu-rep [kernel-substr] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn, path, hdr = None, None, None
agg = {}
seen_fn = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or want not in (fn or ""):
        continue
    if r[0] == "":
        continue          # SASS rows: already summed in the CUDA line row
    n_tail = len(hdr) - 2                      # the source text may itself contain quotes / commas
    src = ",".join(r[1:len(r) - n_tail])
    d = dict(zip(hdr[2:], r[len(r) - n_tail:]))
    key = (fn, path, int(r[0]))
    a = agg.setdefault(key, dict(src=src, samples=0, inst=0, stalls={}))
    a["samples"] += int(d["# Samples"] or 0)
    a["inst"] += int(d["Instructions Executed"] or 0)
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            a["stalls"][k] = a["stalls"].get(k, 0) + int(d[k] or 0)
fns = sorted({k[0] for k in agg})
for f in fns:
    items = [(k, v) for k, v in agg.items() if k[0] == f]
    tot = sum(v["samples"] for _, v in items) or 1
    toti = sum(v["inst"] for _, v in items) or 1
    print(f"== {f[:100]}  samples={tot} warp-inst={toti}")
    for k, v in sorted(items, key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(v["stalls"].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{n[6:]}={c}" for n, c in st if c)
        print(f"  {100*v['samples']/tot:5.1f}% inst {100*v['inst']/toti:5.1f}%  {k[1]}:{k[2]:<4d} {v['src'].strip()[:70]:70s} {sts}")
