import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import ip_mcmc_b200 as M
p = np.load('tests/golden/lorenz_problem_K6_J4.npz'); gs = np.load('tests/golden/lorenz_solves.npz')
for num in ("exact", "fused"):
    for i in range(6):
        T = float(gs[f"case{i}_T"])
        f = M.Lorenz96Moments(6, 4, T, 1, p["prior_means"], p["IC"], numerics=num)
        r = f.batch(gs[f"case{i}_u"].reshape(1, 3), p["IC"].reshape(1, -1))
        acc, rej = r["work"][0].tolist()
        print(num, T, "acc/rej", acc, rej, "ref", int(gs[f"case{i}_n_t"]) - 1, (int(gs[f"case{i}_nfev"]) - 2) // 6 - int(gs[f"case{i}_n_t"]) + 1,
              "max|dG|", np.abs(r["G"].cpu().numpy()[0] - gs[f"case{i}_G"]).max(), flush=True)
