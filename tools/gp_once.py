"""One GP fit + predict(+EI) at n = m = N (default 8192), fixed theta: the GP half of the bench metric, alone (for ncu)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import network_interpretation_imagenet_b200 as nib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = 50
rng = np.random.RandomState(0)
sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(2 * n)]
Z = nib.selection_bits(sels, S)
y = rng.rand(n)
gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=3.0, optimizer=None, query_chunk=8192)
for _ in range(reps):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    gp.fit(Z[:n], y)
    e[1].record()
    mu, var, sd = gp.predict_device(Z[n:])
    nib.expected_improvement_device(mu, sd, float(y.max()), True)
    e[2].record()
    torch.cuda.synchronize()
    print(json.dumps({"n": n, "fit_ms": e[0].elapsed_time(e[1]), "predict_ei_ms": e[1].elapsed_time(e[2])}), flush=True)
