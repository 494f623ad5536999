import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
N, K = 64, 32
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(K + 8, 128, device="cuda", generator=g)   # [k][m]
B = torch.randn(N, K, device="cuda", generator=g)
for off in (0, 1, 2, 3, 4, 5, 8):
    D = torch.full((128, N), -5.0, device="cuda")
    rc = L.mvf_selftest_umma_rows(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, off, 2, st)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("off", off, "EXC", str(e).splitlines()[0]); sys.exit(1)
    ref = A[off:off + K].double().t() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    print("MN-major k-row offset %d rc=%d max err %.4g (ref max %.3g)" % (off, rc, err, ref.abs().max().item()))
