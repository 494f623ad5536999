import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
N, K = 64, 64
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(160, K, device="cuda", generator=g)
B = torch.randn(N, K, device="cuda", generator=g)
for mode in (0, 1):
    for off in (0, 1, 2, 3, 4, 7, 8, 9, 17, 32):
        D = torch.full((128, N), -5.0, device="cuda")
        rc = L.mvf_selftest_umma_rows(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, off, mode, st)
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("mode", mode, "off", off, "EXC", str(e).splitlines()[0]); sys.exit(1)
        ref = A[off:off + 128].double() @ B.double().t()
        err = (D.double() - ref).abs().max().item()
        print("base_off_mode=%d row_off=%2d rc=%d max err %.4g (ref max %.3g)" % (mode, off, rc, err, ref.abs().max().item()))
