import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from poseidon_b200 import _lib as L
    dev = "cuda"
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
    out = {}
    for (rows, C, T) in [(65536, 96, 1024), (16384, 192, 256), (4096, 384, 64), (1024, 768, 16)]:
        dy = torch.randn(rows, C, device=dev); zh = torch.randn(rows, C, device=dev).bfloat16(); rstd = torch.rand(rows, device=dev) + 0.5
        t = torch.rand(rows // T, device=dev); aw = torch.randn(C, device=dev); ab = torch.randn(C, device=dev)
        dz = torch.empty(rows, C, device=dev, dtype=torch.bfloat16); g = [torch.zeros(C, device=dev) for _ in range(5)]
        out[f"{rows}x{C}"] = round(timeit(lambda: L.cln_bwd(dy, zh, rstd, t, aw, ab, dz, False, g[0], g[1], g[2], g[3], g[4], rows, C, T, 0)), 1)
    print(json.dumps({"rpb": os.environ.get("SCOT_CLN_RPB", "default"), **out}))
else:
    for rpb in ["0", "32", "64", "128", "256", "512", "1024"]:
        env = dict(os.environ, SCOT_CLN_RPB=rpb)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
