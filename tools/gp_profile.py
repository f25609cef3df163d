"""Times the individual C-ABI GP calls (CUDA events) at a few sizes: where does fit/predict time go?"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import _lib

lib = _lib.load()
st = _lib.stream_handle()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (M, N, K) in ((4096, 4096, 64), (4096, 4096, 256), (4096, 4096, 4096), (8192, 8192, 64)):
    A = torch.randn(M, K, dtype=torch.float64, device="cuda")
    Bm = torch.randn(K, N, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(M, N, dtype=torch.float64, device="cuda")
    ms = timed(lambda: _lib.check(lib.nib_gp_dgemm_sub(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, st), "dgemm"))
    Cm.zero_()
    _lib.check(lib.nib_gp_dgemm_sub(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, st), "dgemm")
    err = float((Cm + A @ Bm).abs().max())
    print(json.dumps({"dgemm_sub": [M, N, K], "ms": ms, "tflops": 2.0 * M * N * K / (ms * 1e-3) / 1e12, "max_abs_err": err}), flush=True)

for n in (1024, 4096, 8192):
    rng = np.random.RandomState(0)
    S = 50
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(n)]
    Z = torch.from_numpy(nib.selection_bits(sels, S).view(np.int64)).cuda()
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    L = torch.empty_like(K)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    out = {"n": n}
    out["gram_ms"] = timed(lambda: _lib.check(lib.nib_gp_gram_binary(Z.data_ptr(), n, Z.data_ptr(), n, 1, 3.0, 1e-5, K.data_ptr(), n, st), "gram"))

    def chol():
        L.copy_(K)
        _lib.check(lib.nib_gp_cholesky(L.data_ptr(), n, n, info.data_ptr(), st), "chol")
    t_copy = timed(lambda: L.copy_(K))
    out["cholesky_ms"] = timed(chol) - t_copy
    out["cholesky_tflops"] = n ** 3 / 3 / (out["cholesky_ms"] * 1e-3) / 1e12
    v = y.clone()
    out["solve_vec_fwd_ms"] = timed(lambda: _lib.check(lib.nib_gp_trsm(L.data_ptr(), n, n, v.data_ptr(), 1, 1, 0, st), "trsm"))
    out["solve_vec_bwd_ms"] = timed(lambda: _lib.check(lib.nib_gp_trsm(L.data_ptr(), n, n, v.data_ptr(), 1, 1, 1, st), "trsm"))
    out["trsm_n_rhs_ms"] = timed(lambda: _lib.check(lib.nib_gp_trsm(L.data_ptr(), n, n, B.data_ptr(), n, n, 0, st), "trsm"), reps=2)
    out["trsm_tflops"] = float(n) ** 3 / (out["trsm_n_rhs_ms"] * 1e-3) / 1e12
    ref = torch.linalg.cholesky(K)
    out["chol_max_abs_diff_vs_torch"] = float((torch.tril(L) - ref).abs().max())
    print(json.dumps(out), flush=True)
