#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report by SASS basic blocks: runs of
instructions with the same execution count, with their share of executed instructions, of the
stall samples and the opcode mix. Usage: ncu_blocks.py report.ncu-rep [launch_index]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    skip = sys.argv[2] if len(sys.argv) > 2 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen, d = set(), []
    for r in rows[2:]:
        if len(r) <= 10 or r[idx["Address"]] in seen:
            continue
        seen.add(r[idx["Address"]])
        d.append(r)
    print(rows[0][1][:100])
    runs = []
    for r in d:
        n = int(r[idx["Instructions Executed"]] or 0)
        smp = int(r[idx["# Samples"]] or 0)
        tok = r[idx["Source"]].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        if runs and abs(runs[-1]["n"] - n) <= max(1, 0.03 * n):
            x = runs[-1]
            x["k"] += 1; x["s"] += smp; x["ops"][op] += 1; x["tot"] += n
        else:
            runs.append({"n": n, "k": 1, "s": smp, "ops": collections.Counter({op: 1}), "tot": n, "a": r[idx["Address"]][-5:]})
    tot = sum(x["tot"] for x in runs)
    ts = sum(x["s"] for x in runs)
    print(f"instructions {tot}, samples {ts}")
    for x in runs:
        if x["tot"] > 0.004 * tot or x["s"] > 0.01 * ts:
            print(f"{x['a']} exec={x['n']:8d} instrs={x['k']:5d} total={100 * x['tot'] / tot:5.1f}% samples={100 * x['s'] / ts:5.1f}%",
                  dict(x["ops"].most_common(7)))
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    st = {h: sum(int(r[idx[h]] or 0) for r in d) for h in reasons}
    print("stalls:", {k: f"{100 * v / ts:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.03 * ts})


if __name__ == "__main__":
    main()
