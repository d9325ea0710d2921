"""Device-side timing of the LayerNorm forward / backward kernels at the Poseidon-B stage shapes (24 launches cycling over 3
buffer sets > L2, captured in one CUDA graph: the host launch rate does not bound the measurement). `once` = one launch
per shape for ncu captures."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
NSET, NL = 3, 24
once = len(sys.argv) > 1 and sys.argv[1] == "once"
def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(NSET): fn(i)
        torch.cuda.synchronize()
        if once: return 0.0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): fn(i % NSET)
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * NL) * 1e3
B = 64
for (M, C, rps) in [(65536, 96, 1024), (16384, 192, 256), (4096, 384, 64), (1024, 768, 16)]:
    S = [dict(dy=torch.randn(M, C, device=dev), zh=torch.randn(M, C, device=dev).bfloat16(), rstd=torch.rand(M, device=dev) + 0.5,
              dz=torch.empty(M, C, device=dev, dtype=torch.bfloat16), z=torch.randn(M, C, device=dev), res=torch.randn(M, C, device=dev),
              x=torch.empty(M, C, device=dev), xb=torch.empty(M, C, device=dev, dtype=torch.bfloat16)) for _ in range(NSET)]
    t = torch.rand(B, device=dev); aw = torch.randn(C, device=dev); ab = torch.randn(C, device=dev)
    cw = torch.randn(C, device=dev); cb = torch.randn(C, device=dev)
    g = [torch.zeros(C, device=dev) for _ in range(5)]
    bwd = graph_time(lambda i: L.cln_bwd(S[i]["dy"], S[i]["zh"], S[i]["rstd"], t, aw, ab, S[i]["dz"], False, g[0], g[1], g[2], g[3], g[4], M, C, rps))
    fwd = graph_time(lambda i: L.cln_fwd(S[i]["z"], S[i]["res"], t, aw, ab, cw, cb, S[i]["x"], S[i]["xb"], S[i]["zh"], S[i]["rstd"], M, C, rps, 0, 1e-5))
    mb_b = M * C * (4 + 2 + 2) / 1e6; mb_f = M * C * (4 + 4 + 4 + 2 + 2) / 1e6
    print(json.dumps({"shape": [M, C], "bwd_us": round(bwd, 1), "bwd_GBs": round(mb_b / max(bwd, 1e-9) * 1e3), "fwd_us": round(fwd, 1),
                      "fwd_GBs": round(mb_f / max(fwd, 1e-9) * 1e3)}), flush=True)
