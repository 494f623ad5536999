import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import _lib
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
L = _lib.lib()
for mn in (0, 1):
    for N, K in ((16, 32), (64, 64), (128, 96)):
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(128, K, device="cuda", generator=g)
        B = torch.randn(N, K, device="cuda", generator=g)
        D = torch.full((128, N), -5.0, device="cuda")
        Ain = A.t().contiguous() if mn else A
        rc = L.mvf_selftest_umma(Ain.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mn, torch.cuda.current_stream().cuda_stream)
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("mn", mn, N, K, "EXC", str(e)[:80]); sys.exit(1)
        ref = A.double() @ B.double().t()
        err = (D.double() - ref).abs().max().item()
        print("a_mn_major=%d N=%d K=%d rc=%d max err %.4g (ref max %.3g) D[0,:4]=%s ref=%s" % (mn, N, K, rc, err, ref.abs().max().item(), D[0, :4].tolist(), ref[0, :4].float().tolist()))
