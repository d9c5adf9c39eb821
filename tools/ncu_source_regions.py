"""Summarise an `ncu --page source --csv` export: executed warp-instructions and stall samples per address range.

  python tools/ncu_source_regions.py src.csv [bucket_instructions]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    ith = hdr.index("Avg. Threads Executed")
    data = []
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        try:
            data.append((int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), r[isrc], int(r[ismp] or 0), int(r[iex] or 0), float(r[ith] or 0)))
        except ValueError:
            continue
    tot_ex = sum(d[3] for d in data)
    tot_s = sum(d[2] for d in data)
    print("total executed %d, samples %d" % (tot_ex, tot_s))
    base = data[0][0]
    for k in range(0, len(data), bucket):
        chunk = data[k:k + bucket]
        ex = sum(d[3] for d in chunk)
        sm = sum(d[2] for d in chunk)
        ops = {}
        for d in chunk:
            op = d[1].split()[0] if d[1] else "?"
            if op.startswith("@"):
                op = d[1].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + d[3]
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        print("%05x  exec %5.1f%%  samples %5.1f%%  thr %4.1f  %s" % (
            chunk[0][0] - base, 100.0 * ex / tot_ex, 100.0 * sm / max(tot_s, 1),
            sum(d[4] * d[3] for d in chunk) / max(ex, 1), " ".join("%s:%.1f%%" % (o, 100.0 * c / tot_ex) for o, c in top)))


if __name__ == "__main__":
    main()
