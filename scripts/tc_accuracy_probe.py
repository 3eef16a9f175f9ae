"""Where does the 3e-6 error of the tcgen05 split-bf16 GEMM come from?  (run on the GPU box)
Case A: operands exactly representable in bf16 (parts 1,2 are zero) -> any error is the tensor core's
        f32 accumulation (alignment/truncation), not the operand split.
Case B: general fp32 operands (3-part split).
Reports max |err| / max|C| and the mean signed error of entries with |ref| > 0.5 max (a bias means truncation)."""
import ctypes, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import _lib

def run(M, N, K, exact_bf16, positive=False):
    rs = np.random.RandomState(1)
    A = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
    B = torch.from_numpy((rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).cuda()
    if positive:
        A, B = A.abs(), B.abs()
    if exact_bf16:
        A, B = A.bfloat16().float(), B.bfloat16().float()
    C = torch.empty((M, N), dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    parts = []
    for x in (A, B):
        pr = torch.empty((3,) + tuple(x.shape), dtype=torch.bfloat16, device="cuda")
        _lib.call("icrl_split_bf16x3", st, x.numel(), p(x), p(pr), None)
        parts.append(pr)
    _lib.call("icrl_gemm_bf16x3", st, M, N, K, p(parts[0]), p(parts[1]), p(C), N, None, None)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    ref32 = (A @ B.t()).double()          # torch fp32 (cuBLAS, TF32 off by default) for comparison
    d = C.double() - ref
    big = ref.abs() > 0.5 * ref.abs().max()
    rel = (d / ref.abs().max()).abs().max().item()
    bias = ((d[big] / ref[big].abs()) * torch.sign(ref[big])).mean().item()
    rel32 = ((ref32 - ref) / ref.abs().max()).abs().max().item()
    print("M=%d N=%d K=%d exact_bf16=%d positive=%d : max|err|/max|C| = %.3e  signed bias on large entries = %+.3e   (torch fp32: %.3e)"
          % (M, N, K, exact_bf16, positive, rel, bias, rel32))

for K in (64, 128, 512, 2048):
    run(256, 256, K, True)
    run(256, 256, K, True, positive=True)
    run(256, 256, K, False)
    run(256, 256, K, False, positive=True)

# throughput of the standalone split-bf16 GEMM (decode="tc" path): gate GEMM shape of a 4096-row decode step
M, N, K = 4096, 2048, 512
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") / K ** 0.5
C = torch.empty(M, N, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
pa = torch.empty((3, M, K), dtype=torch.bfloat16, device="cuda"); pb = torch.empty((3, N, K), dtype=torch.bfloat16, device="cuda")
_lib.call("icrl_split_bf16x3", st, A.numel(), p(A), p(pa), None)
_lib.call("icrl_split_bf16x3", st, B.numel(), p(B), p(pb), None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(23):
    if i == 3:
        e0.record()
    _lib.call("icrl_gemm_bf16x3", st, M, N, K, p(pa), p(pb), p(C), N, None, None)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print("gemm_bf16x3 %dx%dx%d: %.1f us, algorithmic %.0f TFLOP/s, executed (6 MMAs) %.0f TFLOP/s" % (M, N, K, us, 2.0 * M * N * K / us / 1e6, 12.0 * M * N * K / us / 1e6))
