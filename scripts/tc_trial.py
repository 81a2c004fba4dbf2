"""Bring-up of the tcgen05 linear (csmb_linear_tc) against torch fp64 on the path's shapes; prints error, time, GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csm_mlx_b200 import _lib
dev = torch.device("cuda", 0)
lib = _lib.lib()
torch.manual_seed(0)
shapes = [(16, 3072, 2048), (64, 16384, 2048), (64, 2048, 8192), (10, 2051, 2048), (150, 3072, 2048), (128, 1024, 1024),
          (300, 2048, 2048), (33, 2051, 1024), (64, 16384, 1024), (64, 1024, 8192), (256, 3072, 2048)]
for (R, N, K) in shapes:
    x = torch.randn(R, K, device=dev)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    y = torch.full((R, N), 0.25, device=dev)
    wsb = lib.csmb_linear_tc_workspace_bytes(R, N, K)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    rc = lib.csmb_linear_tc(x.data_ptr(), K, w.data_ptr(), y.data_ptr(), N, R, N, K, 1, ws.data_ptr(), wsb, 0, _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    err_flag = int(ws[:4].view(torch.int32).item())
    ref = x.double() @ w.double().t() + 0.25
    err = (y.double() - ref).abs().max().item()
    # timing (weights rotate so that they come from HBM)
    nw = max(2, int(600e6 / (N * K * 2)))
    ws_list = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(nw)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for wi in ws_list:
            lib.csmb_linear_tc(x.data_ptr(), K, wi.data_ptr(), y.data_ptr(), N, R, N, K, 0, ws.data_ptr(), wsb, 0, _lib.stream_ptr(dev))
    g.replay(); torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nw
    print(f"R={R:4d} N={N:6d} K={K:5d} rc={rc} errflag={err_flag} maxerr={err:.3e} (ref max {ref.abs().max().item():.2f})  {ms*1e3:8.1f} us  {N*K*2/ms/1e6:7.0f} GB/s  {2*R*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
