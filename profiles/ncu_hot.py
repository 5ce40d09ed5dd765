"""Top stall sites of one kernel of an `ncu --set full --import-source on` report (SASS view).

    python profiles/ncu_hot.py rep.ncu-rep [kernel-index (1-based)] [top N]
"""
import csv
import io
import subprocess
import sys


def main(path, kid="1", top="25"):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    allrows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
    k = int(kid) - 1
    rows = allrows[starts[k]:starts[k + 1]]
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[si] or 0) for r in body if len(r) > si)
    print(f"# {rows[0][1] if len(rows[0]) > 1 else ''}  total samples {tot}, {len(body)} SASS instructions")
    ranked = sorted(((int(r[si] or 0), n) for n, r in enumerate(body) if len(r) > si), reverse=True)[:int(top)]
    for s, n in ranked:
        r = body[n]
        why = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{100.0 * s / max(tot, 1):5.1f}%  #{n:4d}  {r[1].strip()[:70]:70s} " + " ".join(f"{w}:{c}" for c, w in why if c))


if __name__ == "__main__":
    main(*sys.argv[1:])
