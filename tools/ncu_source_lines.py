"""Aggregates an ncu source page per CUDA source line: `ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > src.csv`,
then `python tools/ncu_source_lines.py src.csv <kernel name substring> [top N]` prints the share of warp instructions, of stall
samples and the active lanes per instruction of every hot line (inlined callees are attributed to their own file:line)."""
import csv, sys, collections
fn_filter = sys.argv[2]
rows = list(csv.reader(open(sys.argv[1])))
cur_file = cur_fn = None; hdr = None
agg = collections.OrderedDict(); total = 0; tot_samples = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or fn_filter not in (cur_fn or ''): continue
    if len(r) < 10 or r[2] != '-': continue
    try:
        ln = int(r[0]); inst = int(r[7]); samples = int(r[4]); thr = int(r[8])
    except ValueError: continue
    k = (cur_file, ln)
    a = agg.setdefault(k, [0, 0, 0, r[1][:110]])
    a[0] += inst; a[1] += samples; a[2] += thr
    total += inst; tot_samples += samples
print("total warp-instr", total, "samples", tot_samples)
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]
for (f, ln), (inst, smp, thr, src) in top:
    print(f"{100*inst/total:5.1f}% inst {100*smp/max(tot_samples,1):5.1f}% smp  lanes {thr/max(inst,1):4.1f}  {f}:{ln}  {src.strip()}")
