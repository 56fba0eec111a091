"""Accuracy of pf_tc_gemm's split-K mode vs contraction length (diagnostic)."""
import ctypes as C, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pharmacoforge_b200 import _lib
lib = _lib.load()
ws = torch.empty(lib.pf_tc_gemm_workspace_bytes(128, 176, 0) // 4, dtype=torch.float32, device="cuda")
vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
def gemm(A, B, Cm, M, N, K, a_rs, a_cs, b_rs, b_cs):
    _lib.check(lib.pf_tc_gemm(vp(A), vp(B), None, vp(Cm), M, N, K, a_rs, a_cs, b_rs, b_cs, N, 0, vp(ws), ws.numel() * 4,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pf_tc_gemm")
    torch.cuda.synchronize()
for rows, N, K in ((128, 128, 192), (128, 128, 1024), (128, 128, 2048), (128, 161, 4096), (128, 161, 50000), (16, 128, 33333)):
    gen = torch.Generator().manual_seed(1)
    dy = torch.randn(K, rows, generator=gen).cuda(); x = torch.randn(K, N, generator=gen).cuda()
    outs = []
    for _ in range(2):
        dw = torch.full((rows, N), float("nan"), device="cuda"); gemm(dy, x, dw, rows, N, K, 1, rows, N, 1); outs.append(dw)
    ref = dy.double().t() @ x.double()
    ref32 = dy.t() @ x
    err = (outs[0].double() - ref).abs()
    print(rows, N, K, "det", torch.equal(outs[0], outs[1]), "nan", int(torch.isnan(outs[0]).sum()), "max err", float(err.max()),
          "mean signed", float((outs[0].double() - ref).mean()), "torch fp32 err", float((ref32.double() - ref).abs().max()),
          "scale", float(ref.abs().max()))
