"""Summary of an `ncu --page source --csv --print-source sass` export: executed warp instructions and stall samples per
SASS line, grouped into code regions between landmark instructions.  usage: ncu_src_summary.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
ins = [(int(r[col["Instructions Executed"]] or 0), int(r[col["Warp Stall Sampling (All Samples)"]] or 0), r[col["Source"]].strip(), k)
       for k, r in enumerate(data)]
tot_i, tot_s = sum(a for a, *_ in ins), sum(b for _, b, *_ in ins)
print(f"SASS lines {len(ins)}, warp instructions executed {tot_i:,}, stall samples {tot_s:,}")
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: sum(int(r[col[n]] or 0) for r in data) for n in stall_cols}
print("stall reasons:", ", ".join(f"{n[6:]} {100.0 * v / max(tot_s, 1):.1f}%" for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
print(f"\ntop {top} lines by samples:")
for a, b, s, k in sorted(ins, key=lambda t: -t[1])[:top]:
    r = data[k]
    why = max(stall_cols, key=lambda n: int(r[col[n]] or 0))
    print(f"  #{k:5d} samples {b:7d} ({100.0 * b / tot_s:4.1f}%) exec {a:10,d}  {why[6:]:12s} {s[:90]}")
print(f"\nexecuted-instruction profile (runs of equal count):")
run_start, prev = 0, None
for k, (a, b, s, _) in enumerate(ins + [(-1, 0, "", 0)]):
    if a != prev:
        if prev is not None and prev > 0 and (k - run_start) * prev > 0.004 * tot_i:
            smp = sum(x[1] for x in ins[run_start:k])
            print(f"  lines {run_start:5d}-{k - 1:5d} ({k - run_start:4d} instr) x {prev:10,d} = {100.0 * (k - run_start) * prev / tot_i:5.1f}% of instr, "
                  f"{100.0 * smp / tot_s:5.1f}% of samples   first: {ins[run_start][2][:60]}")
        run_start, prev = k, a
