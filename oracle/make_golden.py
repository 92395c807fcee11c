"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py            # all fixtures (~4 min, dominated by Lorenz T_r=500)
    python oracle/make_golden.py burgers    # only the groups named

The fixtures are the pins for oracle/*.py (CPU tests) and for the CUDA engine (GPU tests); they
travel to the GPU box, the reference does not.  Every array below is produced by reference code
(imported through oracle/ref_loader.py); nothing here comes from the restatement.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

PRIOR_MEAN = np.array([1.5, 0.25, -0.5])          # burgers_mcmc.py:63-65
TRUTH = [0.025, -0.025, -0.02]                    # burgers_mcmc.py:29-32
POINTS = [-0.5, -0.25, 0.25, 0.5, 0.75]           # burgers_mcmc.py:50
INTERVAL = 0.1                                    # burgers_mcmc.py:51
NOISE_STD = 0.05                                  # burgers_mcmc.py:55
PRIOR_STD = 0.25                                  # burgers_mcmc.py:66


def burgers_objects(ref, N, domain=(-1, 1), T=1):
    U = ref.utilities
    integ = U.RusanovMCMC(U.BurgersEquation.flux, U.BurgersEquation.flux_prime, domain, N, T)
    meas = U.Measurer(POINTS, INTERVAL, integ.FVM.x[1:-1])
    op = U.FVMObservationOperator(U.PerturbedRiemannIC, PRIOR_MEAN, integ, meas)
    y = meas(integ(U.PerturbedRiemannIC(TRUTH)))            # burgers_mcmc.py:116
    noise = ref.ip_mcmc.GaussianDistribution(np.zeros(5), NOISE_STD ** 2 * np.identity(5))
    pot = ref.ip_mcmc.EvolutionPotential(op, y, noise)
    return integ, meas, op, y, pot


class CountingFVM:
    """Counts FV time steps by wrapping RusanovFVM._step of one instance (no source edits)."""

    def __init__(self, fvm):
        self.n = 0
        orig = fvm._step

        def counted(dt):
            self.n += 1
            return orig(dt)

        fvm._step = counted


def gen_burgers_forward(ref):
    """End states, G(u), Phi(u), FV step counts for several grids and parameter vectors."""
    rng = np.random.default_rng(12345)
    for N in (32, 64, 100, 128, 200, 256):
        integ, meas, op, y, pot = burgers_objects(ref, N)
        counter = CountingFVM(integ.FVM)
        n_u = 6 if N <= 128 else 3
        us = np.zeros((n_u, 3))
        us[1] = np.array(TRUTH) - PRIOR_MEAN          # the truth as a perturbation of the prior mean
        us[2:] = PRIOR_STD * rng.standard_normal((n_u - 2, 3))
        ends, Gs, phis, nfv = [], [], [], []
        for u in us:
            counter.n = 0
            end = integ(ref.utilities.PerturbedRiemannIC(PRIOR_MEAN + u))
            nfv.append(counter.n)
            ends.append(end)
            Gs.append(op(u))
            phis.append(pot(u))
        np.savez(os.path.join(OUT, f"burgers_forward_N{N}.npz"), N=N, u=us, end_state=np.array(ends),
                 G=np.array(Gs), phi=np.array(phis), n_fv=np.array(nfv), y=y, x=integ.FVM.x,
                 dx=integ.FVM.dx, left=meas.left_limits, right=meas.right_limits, dx_meas=meas.dx)
        print("burgers_forward", N, "n_fv", nfv)
    # one large-grid vector (the reference needs ~10 s for it)
    N = 1024
    integ, meas, op, y, pot = burgers_objects(ref, N)
    us = np.array([np.array(TRUTH) - PRIOR_MEAN])
    np.savez(os.path.join(OUT, f"burgers_forward_N{N}.npz"), N=N, u=us,
             end_state=np.array([integ(ref.utilities.PerturbedRiemannIC(PRIOR_MEAN + us[0]))]),
             G=np.array([op(us[0])]), phi=np.array([pot(us[0])]), n_fv=np.array([-1]), y=y,
             x=integ.FVM.x, dx=integ.FVM.dx, left=meas.left_limits, right=meas.right_limits,
             dx_meas=meas.dx)
    print("burgers_forward", N)


class RecordingPotential:
    def __init__(self, pot):
        self.pot = pot
        self.calls = []

    def __call__(self, u):
        val = self.pot(u)
        self.calls.append((np.array(u, dtype=float), float(val)))
        return val

    def exp_minus_potential(self, u):
        return self.pot.exp_minus_potential(u)


def record_chain(ref, sampler_factory, u0, n_steps, name, extra):
    """Run the reference's MCMCSampler.run(u0, n_steps, 0, 1) with a tape RNG."""
    ip = ref.ip_mcmc
    tape = ref_loader.TapeRNG(np.random.default_rng(extra.get("seed", 2)))
    proposer, accepter, rec_pot = sampler_factory(tape)
    sampler = ip.MCMCSampler(proposer, accepter, tape)
    with ref_loader.quiet():
        samples = sampler.run(np.array(u0, dtype=float), n_steps, 0, 1)
    phis_u = np.array([c[1] for c in rec_pot.calls[0::2]])
    phis_v = np.array([c[1] for c in rec_pot.calls[1::2]])
    vs = np.array([c[0] for c in rec_pot.calls[1::2]])
    np.savez(os.path.join(OUT, name), samples=samples, normals=np.array(tape.normals),
             uniforms=np.array(tape.uniforms), phi_u=phis_u, phi_v=phis_v, v=vs,
             calls=accepter.calls, accepts=accepter.accepts, u0=np.array(u0, dtype=float),
             **{k: np.asarray(v) for k, v in extra.items()})
    print(name, "acc", accepter.accepts, "/", accepter.calls)


def gen_burgers_chains(ref, only_n256=False):
    ip = ref.ip_mcmc
    prior = ip.GaussianDistribution(PRIOR_MEAN, PRIOR_STD ** 2 * np.identity(3))

    def pcn(N, beta):
        def f(tape):
            _, _, _, y, pot = burgers_objects(ref, N)
            rec = RecordingPotential(pot)
            return (ip.ConstSteppCNProposer(beta, prior),
                    ip.CountedAccepter(ip.pCNAccepter(rec)), rec)
        return f

    def rw(N, delta):
        def f(tape):
            _, _, _, y, pot = burgers_objects(ref, N)
            rec = RecordingPotential(pot)
            return (ip.ConstStepStandardRWProposer(delta, prior),
                    ip.CountedAccepter(ip.StandardRWAccepter(rec, prior)), rec)
        return f

    # the benchmark configuration (BASELINE.json configs[2]): 256 cells, pCN beta = 0.25 -- from the
    # reference's u_0 = 0 (burgers_mcmc.py:129-134), then continued from its last state with a new tape
    record_chain(ref, pcn(256, 0.25), np.zeros(3), 40, "chain_burgers_pcn_N256.npz",
                 dict(N=256, beta=0.25, seed=2))
    cont = np.load(os.path.join(OUT, "chain_burgers_pcn_N256.npz"))["samples"][-1]   # continue where it ended
    record_chain(ref, pcn(256, 0.25), cont, 150, "chain_burgers_pcn_N256_continued.npz",
                 dict(N=256, beta=0.25, seed=6))
    if only_n256:
        return
    record_chain(ref, pcn(64, 0.25), np.zeros(3), 120, "chain_burgers_pcn_N64.npz",
                 dict(N=64, beta=0.25, seed=2))
    record_chain(ref, pcn(128, 0.15), np.zeros(3), 60, "chain_burgers_pcn_N128.npz",
                 dict(N=128, beta=0.15, seed=3))
    record_chain(ref, rw(64, 0.01125), np.zeros(3), 120, "chain_burgers_rw_N64.npz",
                 dict(N=64, delta=0.01125, seed=2))

    # "next" rows: VarStep RW with the PWLinear schedule of burgers_beta.py:131-147 and
    # ConstrainAccepter(is_valid_IC) of burgers_wasserstein_grid.py:48-56
    class PWLinear:
        def __init__(self, s, e, l):
            self.d_s, self.d_e, self.l = s, e, l
            self.slope = (s - e) / l

        def __call__(self, i):
            if i > self.l:
                return self.d_e
            return self.d_s - self.slope * i

    sched = PWLinear(0.1, 0.001, 50)
    n = 100

    def varstep(tape):
        _, _, _, y, pot = burgers_objects(ref, 64)
        rec = RecordingPotential(pot)
        return (ip.VarStepStandardRWProposer(sched, prior),
                ip.CountedAccepter(ip.StandardRWAccepter(rec, prior)), rec)

    record_chain(ref, varstep, np.zeros(3), n, "chain_burgers_varstep_rw_N64.npz",
                 dict(N=64, seed=4, schedule=[sched(i) for i in range(1, n + 1)]))

    def is_valid(u):
        s = u[2] + PRIOR_MEAN[2]
        return -0.62 < s < -0.38     # a tight box so the constraint actually triggers

    def constrained(tape):
        _, _, _, y, pot = burgers_objects(ref, 64)
        rec = RecordingPotential(pot)
        return (ip.ConstStepStandardRWProposer(0.05, prior),
                ip.CountedAccepter(ip.ConstrainAccepter(ip.StandardRWAccepter(rec, prior), is_valid)),
                rec)

    # with a constraint, potential calls are skipped on violation: record states only
    tape = ref_loader.TapeRNG(np.random.default_rng(5))
    proposer, accepter, rec = constrained(tape)
    sampler = ip.MCMCSampler(proposer, accepter, tape)
    with ref_loader.quiet():
        samples = sampler.run(np.zeros(3), 100, 0, 1)
    np.savez(os.path.join(OUT, "chain_burgers_constrained_rw_N64.npz"), samples=samples,
             normals=np.array(tape.normals), uniforms=np.array(tape.uniforms),
             calls=accepter.calls, accepts=accepter.accepts, u0=np.zeros(3), N=64, delta=0.05,
             lo=-0.62, hi=-0.38, seed=5)
    print("chain_burgers_constrained_rw_N64", accepter.accepts, "/", accepter.calls,
          "uniforms", len(tape.uniforms))


def gen_operator_kats(ref):
    """Values of the reference's operator classes on small inputs (beyond its own unit tests)."""
    ip = ref.ip_mcmc
    rng = np.random.default_rng(7)
    cov = np.array([[2., .5, .1], [.5, 1., .2], [.1, .2, 3.]])
    g = ip.GaussianDistribution(np.zeros(3), cov)
    xs = rng.standard_normal((5, 3))
    logpdf = np.array([g.logpdf(x) for x in xs])
    sqrtcov = np.array([g.apply_sqrt_covariance(x) for x in xs])
    # sample map: same seed -> z and w
    z = np.random.default_rng(11).standard_normal(3)
    w = g.sample(np.random.default_rng(11))
    gd = ip.GaussianDistribution(np.zeros(5), 0.05 ** 2 * np.identity(5))
    devs = rng.standard_normal((4, 5)) * 0.1
    logpdf_diag = np.array([gd.logpdf(x) for x in devs])
    np.savez(os.path.join(OUT, "operator_kats.npz"), cov=cov, xs=xs, logpdf=logpdf, sqrtcov=sqrtcov,
             z=z, w=w, devs=devs, logpdf_diag=logpdf_diag)
    print("operator_kats")


def gen_lorenz(ref):
    L = ref.lorenz
    LM = ref.lorenz_mcmc
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(21)
    # RHS vectors
    rhs_cases = []
    for (K, J) in ((6, 4), (2, 2), (5, 4), (8, 4), (6, 2), (4, 8)):
        F, h, c, b = rng.uniform(1, 12, 4)
        s = rng.standard_normal(K * (J + 1)) * 3
        rhs_cases.append((K, J, F, h, c, b, s, L.Lorenz96(K, J, F, h, c, b)(0, s)))
    np.savez(os.path.join(OUT, "lorenz_rhs.npz"),
             **{f"case{i}_{n}": np.asarray(v) for i, cse in enumerate(rhs_cases)
                for n, v in zip(("K", "J", "F", "h", "c", "b", "state", "rhs"), cse)},
             n_cases=len(rhs_cases))
    print("lorenz_rhs")

    # problem constants (lorenz_mcmc.py:82-139): T_r = 500 data run
    K, J = 6, 4
    theta = np.array([10, 10, 1, 10])
    t0 = time.time()
    Y = LM.run_lorenz96(K, J, theta, 500)
    mf = LM.moment_function(Y, K, J)
    y_data = np.mean(mf, axis=1)
    y_var = np.var(mf, axis=1)
    IC = Y[:, -1].copy()
    print("lorenz data run", Y.shape, f"{time.time() - t0:.1f}s")
    np.savez(os.path.join(OUT, "lorenz_problem_K6_J4.npz"), K=K, J=J, theta=theta, y=y_data, var=y_var,
             IC=IC, n_t=Y.shape[1], r=0.5, prior_means=np.array([12, 8, 9]),
             prior_cov_diag=np.array([10, 1, 10]), T=20, u0=np.array([-1.9, 1.9, 0.9]))

    # solves from the problem IC at several horizons (solve_ivp via the reference's operator)
    prior_means = np.array([12, 8, 9])
    cases = []
    for T in (0.25, 1.0, 2.0, 20.0):
        for u in (np.array([-1.9, 1.9, 0.9]), np.array([-2.5, 1.0, 2.0])):
            op = LM.LorenzObservationOperator(K, J, T, theta[2], prior_means, IC.copy())
            F, h, b = prior_means + u
            sol = solve_ivp(fun=L.Lorenz96(K, J, F, h, theta[2], b), t_span=(0, T), y0=IC.copy(),
                            method='RK45')
            G = op(u)
            cases.append((T, u, G, op.IC.copy(), sol.t.size, sol.nfev, sol.t[:8].copy(), sol.y[:, 1].copy()))
    np.savez(os.path.join(OUT, "lorenz_solves.npz"), n_cases=len(cases),
             **{f"case{i}_{n}": np.asarray(v) for i, cse in enumerate(cases)
                for n, v in zip(("T", "u", "G", "IC_end", "n_t", "nfev", "t_head", "y1"), cse)})
    print("lorenz_solves")

    # a short reference chain with the real problem (pCN beta=0.5, lorenz_mcmc.py:134) at T=2
    ip = ref.ip_mcmc
    noise = ip.GaussianDistribution(np.zeros(30), 0.5 ** 2 * np.diag(y_var))
    prior = ip.GaussianDistribution(np.zeros(3), np.diag([10, 1, 10]))
    for (name, T, n) in (("chain_lorenz_pcn_T2.npz", 2, 40), ("chain_lorenz_pcn_T20.npz", 20, 6)):
        def fac(tape, T=T):
            op = LM.LorenzObservationOperator(K, J, T, theta[2], prior_means, IC.copy())
            rec = RecordingPotential(ip.EvolutionPotential(op, y_data, noise))
            return (ip.ConstSteppCNProposer(0.5, prior), ip.CountedAccepter(ip.pCNAccepter(rec)), rec)
        record_chain(ref, fac, np.array([-1.9, 1.9, 0.9]), n, name, dict(T=T, beta=0.5, seed=1))


def gen_burgers_kl(ref):
    """EXTENSION fixture: the reference's RusanovFVM / Measurer / EvolutionPotential driven by an
    initial condition of the form Riemann(x) + sum_k a_k phi_k(x) (the IC callable below is ours,
    everything downstream of it is reference code)."""
    U = ref.utilities
    rng = np.random.default_rng(99)
    for N, m in ((64, 4), (128, 8)):
        integ = U.RusanovMCMC(U.BurgersEquation.flux, U.BurgersEquation.flux_prime, (-1, 1), N, 1)
        x = integ.FVM.x
        meas = U.Measurer(POINTS, INTERVAL, x[1:-1])
        k = np.arange(1, m + 1, dtype=np.float64)[:, None]
        basis = np.sin(k * np.pi * (x[None, :] - (-1)) / 2)
        index = {float(xv): i for i, xv in enumerate(x)}

        class RiemannKLIC:
            def __init__(self, params):
                self.base = U.PerturbedRiemannIC(params[:3])
                self.a = params[3:]

            def __call__(self, xv):
                val = self.base(xv)
                i = index[float(xv)]
                for kk in range(m):
                    val = val + self.a[kk] * basis[kk, i]
                return val

        mean = np.concatenate([PRIOR_MEAN, np.zeros(m)])
        truth = np.concatenate([TRUTH, 0.05 * rng.standard_normal(m) / np.arange(1, m + 1)])
        y = meas(integ(RiemannKLIC(truth)))
        noise = ref.ip_mcmc.GaussianDistribution(np.zeros(5), NOISE_STD ** 2 * np.identity(5))
        op = U.FVMObservationOperator(RiemannKLIC, mean, integ, meas)
        pot = ref.ip_mcmc.EvolutionPotential(op, y, noise)
        us = np.zeros((4, 3 + m))
        us[1] = truth - mean
        us[2:, :3] = PRIOR_STD * rng.standard_normal((2, 3))
        us[2:, 3:] = 0.1 * rng.standard_normal((2, m)) / np.arange(1, m + 1)
        ends = [integ(RiemannKLIC(mean + u)) for u in us]
        np.savez(os.path.join(OUT, f"burgers_kl_N{N}_m{m}.npz"), N=N, m=m, u=us, truth=truth, y=y, basis=basis,
                 end_state=np.array(ends), G=np.array([op(u) for u in us]), phi=np.array([pot(u) for u in us]))
        print("burgers_kl", N, m)


def gen_tau0(ref):
    """tau_0 of the reference (utilities.uncorrelated_sample_spacing, utilities.py:169-186, built on
    helpers.autocorrelation, helpers.py:41-54).  helpers.py cannot be imported (it imports POT and runs a
    test at import), so the ONE function is compiled from its source file in place (nothing is copied)."""
    import ast
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, "report", "scripts", "helpers.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "autocorrelation"]
    ns = dict(np=np, MCMCSampler=ref.ip_mcmc.MCMCSampler)
    exec(compile(ast.Module(body=fn, type_ignores=[]), "helpers.py", "exec"), ns)
    ref.utilities.autocorrelation = ns["autocorrelation"]
    rng = np.random.default_rng(12)
    out = {}
    for i, (n, rho) in enumerate([(1500, 0.9), (1500, 0.5), (400, 0.97), (25, 0.99), (3000, 0.0), (12, 0.9)]):
        x = np.empty((3, n))
        e = rng.standard_normal((3, n))
        x[:, 0] = e[:, 0]
        for k in range(1, n):
            x[:, k] = rho * x[:, k - 1] + np.sqrt(1 - rho ** 2) * e[:, k]
        out[f"case{i}_x"] = x
        out[f"case{i}_tau0"] = ref.utilities.uncorrelated_sample_spacing(x)
        out[f"case{i}_ac20"] = ns["autocorrelation"](x, 20) if n >= 20 else np.zeros((3, 20))
        print("tau0 case", i, n, rho, "->", out[f"case{i}_tau0"])
    out["n_cases"] = 6
    np.savez(os.path.join(OUT, "tau0_reference.npz"), **out)


GROUPS = dict(burgers=gen_burgers_forward, kl=gen_burgers_kl, chains=gen_burgers_chains, kats=gen_operator_kats,
              lorenz=gen_lorenz, tau0=gen_tau0, chains256=lambda ref: gen_burgers_chains(ref, only_n256=True))

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    which = sys.argv[1:] or list(GROUPS)
    for g in which:
        t0 = time.time()
        GROUPS[g](ref)
        print(f"[{g}] done in {time.time() - t0:.1f}s")
