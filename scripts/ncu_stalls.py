"""Where do the warps of a kernel wait? Reads an .ncu-rep (captured with `--set full --import-source on`) offline and
prints, per kernel: the stall-reason totals of the warp-state samples, the executed-instruction mix (opcode histogram
weighted by execution counts) and the SASS lines with the most samples.

    python scripts/ncu_stalls.py gpurun_out/r01_gemm_async_full.ncu-rep [top=12] [kernel-substring]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for k in kernels:
        if k["name"] in seen or want not in k["name"] or "Source" not in (k["hdr"] or []):
            continue
        seen.add(k["name"])
        h = k["hdr"]
        isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stalls = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[isamp] or 0) for r in k["rows"]) or 1
        print("=====", k["name"][:110])
        st = {c: sum(int(r[i] or 0) for r in k["rows"]) for i, c in stalls}
        print("samples", tot, {c: v for c, v in sorted(st.items(), key=lambda kv: -kv[1]) if v})
        ops = collections.Counter()
        for r in k["rows"]:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
            if m:
                ops[m.group(2)] += int(r[iex] or 0)
        n = sum(ops.values()) or 1
        print("executed warp-instructions", n, [(o, round(100 * c / n, 1)) for o, c in ops.most_common(14)])
        for r in sorted(k["rows"], key=lambda r: -int(r[isamp] or 0))[:top]:
            dom = max(stalls, key=lambda ic: int(r[ic[0]] or 0))[1]
            print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  executed={r[iex]:>8s} {dom:22s} {r[isrc][:90]}")


if __name__ == "__main__":
    main()
