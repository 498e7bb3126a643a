set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_posterior.py tests/test_gpu_fit.py -q -m gpu 2>&1 | tail -4
timeout 600 python tools/bench_configs.py c3 2>&1 | tail -2
python - <<'PY'
import time, numpy as np, gpax_b200
rng = np.random.default_rng(0)
N, d = 16384, 3
X = rng.uniform(0, 1, (N, d)); y = rng.standard_normal(N)
params = {"k_length": np.full(d, 0.3), "k_scale": 1.0, "noise": 0.1}
m = gpax_b200.viGP(d, "RBF"); m.X_train, m.y_train = X, y
m.predict(0, rng.uniform(0, 1, (64, d)), params)
for P in (100, 1000, 4096):
    Xn = rng.uniform(0, 1, (P, d))
    m.predict(0, Xn, params)
    t0 = time.perf_counter()
    for _ in range(3): m.predict(0, Xn, params)
    print(f"reuse solve P={P}: {(time.perf_counter()-t0)/3*1e3:.2f} ms per call", flush=True)
PY
