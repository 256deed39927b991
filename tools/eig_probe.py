import sys, time, torch
sys.path.insert(0, "/root/repo")
from dftbplus_b200.api import HotPath, new_dense
hp = HotPath(0); hp.use_torch_stream()
def run(n, cplx, reps=1):
    dt = torch.complex128 if cplx else torch.float64
    A = torch.randn(n, n, dtype=dt, device="cuda")
    H = new_dense(n, dtype=dt); S = new_dense(n, dtype=dt)
    H.copy_(A + A.conj().t()); 
    S.copy_(torch.eye(n, dtype=dt, device="cuda") * 2 + 0.01 * (A + A.conj().t()) / n ** 0.5)
    del A
    torch.cuda.synchronize()
    for r in range(reps + 1):
        Hc = new_dense(n, dtype=dt); Sc = new_dense(n, dtype=dt); Hc.copy_(H); Sc.copy_(S)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ev = hp.diagDenseMtx(Hc, Sc)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        print("n", n, "cplx", cplx, "rep", r, "s", round(t, 3), "mem GB", round(torch.cuda.max_memory_allocated() / 1e9, 1), flush=True)
        del Hc, Sc
import os
if "--big" not in sys.argv:
    run(2048, True, 2)
    run(4092, False, 2)
run(31744, False, 0)
