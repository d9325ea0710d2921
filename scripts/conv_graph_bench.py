"""Device-side timing (CUDA graph of 12 launches over 3 buffer sets) of the ConvNeXt depthwise 7x7 kernels at the Poseidon-B
skip-connection shapes: forward, backward (data gradient + weight gradient launches together)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
NSET, NL = 3, 12
def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(NSET): fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): fn(i % NSET)
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * NL) * 1e3
B = 64
for res, C in [(32, 96), (16, 192), (8, 384)]:
    S = [dict(x=torch.randn(B, res, res, C, device=dev), out=torch.empty(B, res, res, C, device=dev), dout=torch.randn(B, res, res, C, device=dev),
              gin=torch.randn(B, res, res, C, device=dev), gout=torch.empty(B, res, res, C, device=dev)) for _ in range(NSET)]
    w = torch.randn(C, 49, device=dev); b = torch.randn(C, device=dev); gw = torch.zeros(C, 49, device=dev)
    fwd = graph_time(lambda i: L.glue("scot_convnext_dwconv7_fwd", S[i]["x"], w, b, S[i]["out"], B, res, C))
    bwd = graph_time(lambda i: L.glue("scot_convnext_dwconv7_bwd", S[i]["x"], w, S[i]["dout"], S[i]["gin"], S[i]["gout"], gw, B, res, C))
    print(json.dumps({"res": res, "C": C, "fwd_us": round(fwd, 1), "bwd_data_plus_wgrad_us": round(bwd, 1)}), flush=True)
