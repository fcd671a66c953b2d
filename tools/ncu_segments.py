"""Aggregates the warp-stall samples of an `ncu --page source --csv` export into the code segments between
barriers -- shows which phase of a shared-memory FFT kernel the time goes to.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv; python tools/ncu_segments.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isamp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
names = ["stall_barrier", "stall_lg", "stall_long_sb", "stall_math", "stall_mio", "stall_not_selected", "stall_selected",
         "stall_short_sb", "stall_wait", "stall_dispatch", "stall_no_inst"]
idx = {n: hdr.index(n) for n in names}


def num(r, i):
    try:
        return int(r[i])
    except (ValueError, IndexError):
        return 0


body = [r for r in rows[2:] if len(r) > isamp]
tot = sum(num(r, isamp) for r in body)
print("total samples", tot, "instructions", len(body))
acc, start, st = 0, 0, {n: 0 for n in names}
ops = {}
for i, r in enumerate(body):
    acc += num(r, isamp)
    for n in names:
        st[n] += num(r, idx[n])
    w = r[ia].split()
    op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "")).split(".")[0]
    ops[op] = ops.get(op, 0) + 1
    if "BAR.SYNC" in r[ia] or i == len(body) - 1:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        mix = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"instr {start}-{i}: {acc} ({100 * acc / tot:.1f}%)  " + " ".join(f"{k[6:]}={v}" for k, v in top) + "  | " + mix)
        acc, start, st, ops = 0, i + 1, {n: 0 for n in names}, {}
