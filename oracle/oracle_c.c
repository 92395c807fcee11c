/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's hot path, built with
 *   gcc -O2 -ffp-contract=off -pthread -shared -fPIC        (oracle/c_oracle.py:build -> oracle/_build/liboracle_c.so)
 * so that statistically meaningful CPU chains (tens of thousands of chain-steps at the bench grids)
 * finish in seconds.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it; nothing
 * under ip_mcmc_b200/ does.
 *
 * Restates (paths relative to the reference root), in the reference's floating-point operation order:
 *   RusanovFVM.integrate/_step/_flux/_cfl/_apply_bc   report/scripts/burgers/rusanov.py:31-109
 *   BurgersEquation.flux/flux_prime                   report/scripts/burgers/utilities.py:112-119
 *   PerturbedRiemannIC                                report/scripts/burgers/utilities.py:44-62
 *   Measurer.__call__                                 report/scripts/burgers/utilities.py:100-109
 *   FVMObservationOperator.__call__                   report/scripts/burgers/utilities.py:40-41
 *   Lorenz96.__call__                                 report/scripts/lorenz.py:44-101
 *   moment_function / LorenzObservationOperator       report/scripts/lorenz_mcmc.py:17-71
 *   scipy RK45 (solve_ivp defaults)                   scipy/integrate/_ivp/rk.py:14-175, common.py:63-140
 *   EvolutionPotential.__call__ (diagonal noise)      ip_mcmc/ip_mcmc/potential.py:53-54
 *   ConstStepStandardRW/pCN proposers                 ip_mcmc/ip_mcmc/proposer.py:14-30, 59-82
 *   ProbabilisticAccepter, StandardRW/pCN accepters   ip_mcmc/ip_mcmc/accepter.py:58-122
 *   MCMCSampler._step                                 ip_mcmc/ip_mcmc/sampler.py:35-41
 *
 * Pinning (tests/test_oracle_c.py): the Burgers path is BIT-IDENTICAL to oracle/burgers_np.py -- itself
 * bit-identical to the live reference on tests/golden/burgers_*.npz -- and replays the reference's recorded
 * chains (tests/golden/chain_burgers_*.npz) state by state.  The Lorenz path agrees with oracle/lorenz_np.py
 * (bit-identical to scipy on this stack) to rounding per RK attempt: scipy's stage sums go through BLAS
 * (np.dot), whose summation order / FMA use is not part of any specification, so bit-identity is not
 * defined there; the right-hand side itself is bit-identical.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * shared helpers
 * ---------------------------------------------------------------------------------------------- */
static double max_nan(double a, double b) { /* np.maximum / np.max: NaN propagates */
    if (isnan(a)) return a;
    if (isnan(b)) return b;
    return a > b ? a : b;
}

/* numpy DOUBLE_pairwise_sum (loops_utils.h.src) of a[0..n) */
static double pairwise_sum(const double *a, long n) {
    if (n < 8) {
        double res = -0.0;
        for (long i = 0; i < n; ++i) res = res + a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        long i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] = r[k] + a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res = res + a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

/* Phi = 0.5*((rank*log2pi + log_pdet) + sum_i r_i^2), r_i = (y - G)[perm_i] * scale_i
 * (scipy _multivariate.py:585-591 for a diagonal covariance; potential.py:53-54) */
typedef struct {
    int q;
    const double *y;
    const int *perm;
    const double *scale;
    double log_const;
} orc_potential;

static double potential_from_G(const orc_potential *P, const double *G) {
    double r2[64];
    for (int i = 0; i < P->q; ++i) {
        const int j = P->perm[i];
        const double r = (P->y[j] - G[j]) * P->scale[i];
        r2[i] = r * r;
    }
    const double maha = pairwise_sum(r2, P->q);
    return 0.5 * (P->log_const + maha);
}

/* ------------------------------------------------------------------------------------------------
 * Burgers
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int N;              /* interior cells */
    long max_steps;     /* <= 0: unlimited (as the reference) */
    double T, dx, dx_meas;
    const double *x;    /* [N+2] cell centres incl. ghosts */
    const int *left, *right; /* [q] windows into the interior array */
    const double *prior_mean; /* [3] */
    orc_potential pot;
} orc_burgers;

/* rusanov.py:92-96 with f(w) = .5*w*w, f'(w) = w */
static inline double rusanov_flux(double ul, double ur) {
    const double fl = .5 * ul * ul, fr = .5 * ur * ur;
    const double flux_average = 0.5 * (fl + fr);
    const double speed = max_nan(fabs(ul), fabs(ur));
    return flux_average - 0.5 * speed * (ur - ul);
}

/* dudt[1..N] = (F_{i+1/2} - F_{i-1/2}) / (-dx), ghosts 0 (rusanov.py:76-90) */
static void rate_of_change(const double *u, int N, double dx, double *F, double *dudt) {
    for (int i = 0; i < N + 1; ++i) F[i] = rusanov_flux(u[i], u[i + 1]);
    dudt[0] = 0.0 / -dx;
    dudt[N + 1] = 0.0 / -dx;
    for (int i = 1; i <= N; ++i) dudt[i] = (F[i] - F[i - 1]) / -dx;
}

/* integrate PerturbedRiemannIC(params) to t >= T; returns the number of FV steps, end state in out[N] */
long orc_burgers_solve(const orc_burgers *B, const double *params, double *out, double *t_end) {
    const int N = B->N, M = N + 2;
    double *u = (double *)malloc(sizeof(double) * M * 4), *us = u + M, *F = us + M, *dudt = F + M;
    const double left = 1 + params[0], right = params[1], jump = params[2];
    for (int i = 0; i < M; ++i) u[i] = (B->x[i] < jump) ? left : right;   /* rusanov.py:32, utilities.py:59-62 */
    double t = 0;
    long n = 0;
    const double dx = B->dx;
    while (t < B->T) {                                                     /* rusanov.py:40-45 */
        if (B->max_steps > 0 && n >= B->max_steps) break;
        double m = fabs(u[1]);
        for (int i = 2; i <= N; ++i) m = max_nan(m, fabs(u[i]));           /* rusanov.py:102-109 */
        const double dt = 0.5 * dx / m;
        t += dt;
        /* SSPRK2, rusanov.py:62-74 */
        rate_of_change(u, N, dx, F, dudt);
        for (int i = 0; i < M; ++i) us[i] = u[i] + dt * dudt[i];
        us[0] = us[1];
        us[M - 1] = us[M - 2];
        rate_of_change(us, N, dx, F, dudt);
        for (int i = 0; i < M; ++i) us[i] += dt * dudt[i];
        for (int i = 0; i < M; ++i) u[i] = (u[i] + us[i]) / 2;
        u[0] = u[1];
        u[M - 1] = u[M - 2];
        ++n;
    }
    memcpy(out, u + 1, sizeof(double) * N);
    if (t_end) *t_end = t;
    free(u);
    return n;
}

/* Measurer.__call__ (utilities.py:100-109): 10 * trapz(values[l:r], dx), numpy's evaluation order */
static void burgers_measure(const orc_burgers *B, const double *state, double *G) {
    double terms[4096];
    for (int i = 0; i < B->pot.q; ++i) {
        const int l = B->left[i], r = B->right[i], nt = r - l - 1;
        double s = 0.0;
        if (nt >= 1) {
            for (int j = 0; j < nt; ++j) terms[j] = B->dx_meas * (state[l + j + 1] + state[l + j]) / 2.0;
            s = pairwise_sum(terms, nt);
        }
        G[i] = 10 * s;
    }
}

/* G(u) = meas(integrate(IC(prior_mean + u))) and Phi(u); returns Phi */
double orc_burgers_phi(const orc_burgers *B, const double *u, double *G_out, double *state_out, long *n_fv) {
    double params[3], G[64];
    for (int i = 0; i < 3; ++i) params[i] = B->prior_mean[i] + u[i];
    double *state = (double *)malloc(sizeof(double) * B->N);
    double t_end;
    const long n = orc_burgers_solve(B, params, state, &t_end);
    burgers_measure(B, state, G);
    if (G_out) memcpy(G_out, G, sizeof(double) * B->pot.q);
    if (state_out) memcpy(state_out, state, sizeof(double) * B->N);
    if (n_fv) *n_fv = n;
    free(state);
    const double phi = potential_from_G(&B->pot, G);
    return (t_end < B->T) ? NAN : phi;   /* capped solve: the engine's convention */
}

/* ------------------------------------------------------------------------------------------------
 * Lorenz-96 + RK45
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int K, J;
    long max_attempts;
    double T, c, rtol, atol;
    const double *prior_mean; /* [3]: F, h, b */
    orc_potential pot;
} orc_lorenz;

#define LMAX 512 /* K*(J+1) */

/* lorenz.py:44-101; state = [X_0..X_{K-1}, Y_{0,0..J-1}, ...] */
void orc_lorenz_rhs(int K, int J, double F, double h, double c, double b, const double *s, double *out) {
    const double *X = s, *Y = s + K;
    const double hc = h * c, hJ = h / J;
    for (int k = 0; k < K; ++k) {
        const double Xm1 = X[(k + K - 1) % K], Xm2 = X[(k + 2 * K - 2) % K], Xp1 = X[(k + 1) % K];
        double o = X[k] * -1;
        o -= Xm1 * Xm2 - Xm1 * Xp1;
        o += F;
        if (J > 0) {
            const double *Yk = Y + (long)k * J;
            const double sum = pairwise_sum(Yk, J);
            o -= hc * (sum / J);
            const double hx = hJ * X[k];
            for (int j = 0; j < J; ++j) {
                const double Yp1 = Yk[(j + 1) % J], Yp2 = Yk[(j + 2) % J], Ym1 = Yk[(j + J - 1) % J];
                double y = Yk[j] * -1;
                y -= b * (Yp1 * Yp2 - Ym1 * Yp1);
                y += hx;
                y *= c;
                out[K + (long)k * J + j] = y;
            }
        }
        out[k] = o;
    }
}

static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

static double rms_norm(const double *x, int n) { /* common.py:63-65 */
    double ss = 0.0;
    for (int i = 0; i < n; ++i) ss += x[i] * x[i];
    return sqrt(ss) / sqrt((double)n);
}

typedef struct {
    int K, J, n;
    double F, h, c, b;
} lorenz_fun;

static void lfun(const lorenz_fun *f, const double *y, double *out) { orc_lorenz_rhs(f->K, f->J, f->F, f->h, f->c, f->b, y, out); }

/* one Dormand-Prince attempt (rk.py:14-72, 111-116): Kst[0] = f on entry; returns error_norm */
double orc_rk45_attempt(int K, int J, const double *theta4, const double *y, const double *f, double h, double rtol,
                        double atol, double *y_new, double *f_new) {
    const lorenz_fun fn = {K, J, K * (J + 1), theta4[0], theta4[1], theta4[2], theta4[3]};
    const int n = fn.n;
    double Kst[7][LMAX], ys[LMAX], tmp[LMAX];
    memcpy(Kst[0], f, sizeof(double) * n);
    for (int s = 1; s < 6; ++s) {
        for (int i = 0; i < n; ++i) {
            double dy = 0.0;
            for (int j = 0; j < s; ++j) dy += Kst[j][i] * RK_A[s][j];
            ys[i] = y[i] + dy * h;
        }
        lfun(&fn, ys, Kst[s]);
    }
    for (int i = 0; i < n; ++i) {
        double d = 0.0;
        for (int j = 0; j < 6; ++j) d += Kst[j][i] * RK_B[j];
        y_new[i] = y[i] + h * d;
    }
    lfun(&fn, y_new, Kst[6]);
    memcpy(f_new, Kst[6], sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        double e = 0.0;
        for (int j = 0; j < 7; ++j) e += Kst[j][i] * RK_E[j];
        const double scale = atol + max_nan(fabs(y[i]), fabs(y_new[i])) * rtol;
        tmp[i] = e * h / scale;
    }
    return rms_norm(tmp, n);
}

/* LorenzObservationOperator.__call__ (lorenz_mcmc.py:55-68): IC in/out; G[5K] out; returns Phi */
double orc_lorenz_phi(const orc_lorenz *L, const double *u, double *IC, double *G_out, long *n_acc_out, long *n_rej_out) {
    const int K = L->K, J = L->J, n = K * (J + 1);
    const double theta[4] = {L->prior_mean[0] + u[0], L->prior_mean[1] + u[1], L->c, L->prior_mean[2] + u[2]};
    const lorenz_fun fn = {K, J, n, theta[0], theta[1], theta[2], theta[3]};
    double y[LMAX], f[LMAX], y_new[LMAX], f_new[LMAX], tmp[LMAX], scale[LMAX];
    double msum[5 * 64];
    memcpy(y, IC, sizeof(double) * n);
    for (int i = 0; i < 5 * K; ++i) msum[i] = 0.0;
    long n_t = 0, n_acc = 0, n_rej = 0;
#define ADD_MOMENTS()                                         \
    do {                                                      \
        for (int k = 0; k < K; ++k) {                         \
            const double X = y[k], Y0 = y[K + (long)k * J];   \
            msum[k] += X;                                     \
            msum[K + k] += Y0;                                \
            msum[2 * K + k] += X * X;                         \
            msum[3 * K + k] += X * Y0;                        \
            msum[4 * K + k] += Y0 * Y0;                       \
        }                                                     \
        ++n_t;                                                \
    } while (0)
    double t = 0.0;
    const double t_bound = L->T;
    lfun(&fn, y, f);
    ADD_MOMENTS();
    /* select_initial_step (common.py:68-140), order 4 */
    double h_abs;
    {
        for (int i = 0; i < n; ++i) scale[i] = L->atol + fabs(y[i]) * L->rtol;
        for (int i = 0; i < n; ++i) tmp[i] = y[i] / scale[i];
        const double d0 = rms_norm(tmp, n);
        for (int i = 0; i < n; ++i) tmp[i] = f[i] / scale[i];
        const double d1 = rms_norm(tmp, n);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (t_bound < h0) h0 = t_bound;
        for (int i = 0; i < n; ++i) y_new[i] = y[i] + h0 * 1 * f[i];
        lfun(&fn, y_new, f_new);
        for (int i = 0; i < n; ++i) tmp[i] = (f_new[i] - f[i]) / scale[i];
        const double d2 = rms_norm(tmp, n) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), t_bound);
    }
    int status = 0;
    while (t != t_bound && status == 0) {                      /* rk.py:118-175 */
        const double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int step_accepted = 0, step_rejected = 0;
        double t_new = t;
        while (!step_accepted) {
            if (h_abs < min_step || (L->max_attempts > 0 && n_acc + n_rej >= L->max_attempts)) {
                status = -1;
                break;
            }
            double h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            const double err = orc_rk45_attempt(K, J, theta, y, f, h, L->rtol, L->atol, y_new, f_new);
            if (err < 1) {
                double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
                if (step_rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                step_accepted = 1;
                ++n_acc;
            } else {
                h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));   /* NaN error norm: max(0.2, nan) = 0.2 in Python */
                step_rejected = 1;
                ++n_rej;
            }
        }
        if (status != 0) break;
        t = t_new;
        memcpy(y, y_new, sizeof(double) * n);
        memcpy(f, f_new, sizeof(double) * n);
        ADD_MOMENTS();
    }
#undef ADD_MOMENTS
    memcpy(IC, y, sizeof(double) * n);                          /* lorenz_mcmc.py:66 */
    double G[5 * 64];
    for (int i = 0; i < 5 * K; ++i) G[i] = msum[i] / (double)n_t;
    if (G_out) memcpy(G_out, G, sizeof(double) * 5 * K);
    if (n_acc_out) *n_acc_out = n_acc;
    if (n_rej_out) *n_rej_out = n_rej;
    return potential_from_G(&L->pot, G);
}

/* ------------------------------------------------------------------------------------------------
 * Metropolis chains with injected noise (the reference's MockRNG seam, test_utilities.py:11-26)
 * ---------------------------------------------------------------------------------------------- */
enum { ORC_RW = 0, ORC_PCN = 1 };

/* accepter.py:104-106: .5 * norm(L @ w)**2 */
static double prior_regulariser(const double *L, const double *w, int d) {
    double ss = 0.0;
    for (int i = 0; i < d; ++i) {
        double y = 0.0;
        for (int j = 0; j < d; ++j) y += L[i * d + j] * w[j];
        ss += y * y;
    }
    const double nrm = sqrt(ss);
    return .5 * (nrm * nrm);
}

/* n_chains independent chains x n_steps (sampler.py:35-41), POSIX threads over chains.
 *   model: 0 Burgers (B != NULL), 1 Lorenz (L != NULL; IC [n_chains, nvar] in/out)
 *   normals [n_chains, n_steps, 3] at the w level (proposal draws N(0,C)), uniforms [n_chains, n_steps]
 *   recompute_phi_u: evaluate Phi(u) again every step before Phi(v) (accepter.py:99-100,121-122); the
 *                    reference always does; mandatory for the stateful Lorenz operator
 *   outputs (may be NULL): states [n_chains, n_steps, 3] (state after each step), phi_v [n_chains, n_steps],
 *                    accepted [n_chains, n_steps] (uint8), work [n_chains, 2] (Burgers: FV steps, solves;
 *                    Lorenz: accepted, rejected RK attempts) */
typedef struct {
    const orc_burgers *B;
    const orc_lorenz *L;
    int n_chains, proposer, accepter, recompute_phi_u;
    long n_steps;
    double ca, cb;
    const double *prior_chol, *u0, *normals, *uniforms;
    double *IC, *states, *phi_v_out;
    unsigned char *accepted_out;
    long *work;
    int next; /* next chain to take (atomic) */
} chain_job;

static void run_one_chain(const chain_job *Jb, int c) {
    const orc_burgers *B = Jb->B;
    const orc_lorenz *L = Jb->L;
    const int d = 3;
    const int nvar = L ? L->K * (L->J + 1) : 0;
    const long n_steps = Jb->n_steps;
    double u[3], v[3], phi_u = NAN, phi_v;
    long wa = 0, wb = 0;
    for (int i = 0; i < d; ++i) u[i] = Jb->u0[(long)c * d + i];
    double *ic = L ? Jb->IC + (long)c * nvar : NULL;
    for (long s = 0; s < n_steps; ++s) {
        const double *w = Jb->normals + ((long)c * n_steps + s) * d;
        for (int i = 0; i < d; ++i) v[i] = (Jb->proposer == ORC_RW) ? u[i] + Jb->cb * w[i] : Jb->ca * u[i] + Jb->cb * w[i];
        long a = 0, b = 0;
        if (isnan(phi_u) || Jb->recompute_phi_u) {
            if (B) { phi_u = orc_burgers_phi(B, u, NULL, NULL, &a); wa += a; wb += 1; }
            else { phi_u = orc_lorenz_phi(L, u, ic, NULL, &a, &b); wa += a; wb += b; }
        }
        if (B) { phi_v = orc_burgers_phi(B, v, NULL, NULL, &a); wa += a; wb += 1; }
        else { phi_v = orc_lorenz_phi(L, v, ic, NULL, &a, &b); wa += a; wb += b; }
        double acc_p;
        if (Jb->accepter == ORC_RW)
            acc_p = exp((phi_u + prior_regulariser(Jb->prior_chol, u, d)) - (phi_v + prior_regulariser(Jb->prior_chol, v, d)));
        else
            acc_p = exp(phi_u - phi_v);
        const int acc = acc_p > Jb->uniforms[(long)c * n_steps + s];   /* strict, un-clipped; NaN rejects */
        if (acc) {
            for (int i = 0; i < d; ++i) u[i] = v[i];
            phi_u = phi_v;
        }
        if (Jb->states)
            for (int i = 0; i < d; ++i) Jb->states[((long)c * n_steps + s) * d + i] = u[i];
        if (Jb->phi_v_out) Jb->phi_v_out[(long)c * n_steps + s] = phi_v;
        if (Jb->accepted_out) Jb->accepted_out[(long)c * n_steps + s] = (unsigned char)acc;
    }
    if (Jb->work) {
        Jb->work[2 * c] = wa;
        Jb->work[2 * c + 1] = wb;
    }
}

static void *chain_worker(void *arg) {
    chain_job *Jb = (chain_job *)arg;
    for (;;) {
        const int c = __atomic_fetch_add(&Jb->next, 1, __ATOMIC_RELAXED);
        if (c >= Jb->n_chains) break;
        run_one_chain(Jb, c);
    }
    return NULL;
}

int orc_run_chains(const orc_burgers *B, const orc_lorenz *L, int n_chains, long n_steps, int proposer, int accepter,
                   double step, const double *prior_chol, const double *u0, double *IC, const double *normals,
                   const double *uniforms, int recompute_phi_u, double *states, double *phi_v_out,
                   unsigned char *accepted_out, long *work, int n_threads) {
    chain_job Jb;
    memset(&Jb, 0, sizeof Jb);
    Jb.B = B; Jb.L = L; Jb.n_chains = n_chains; Jb.n_steps = n_steps; Jb.proposer = proposer; Jb.accepter = accepter;
    Jb.recompute_phi_u = L ? 1 : recompute_phi_u;
    if (proposer == ORC_RW) {
        Jb.ca = 1.0;
        Jb.cb = sqrt(2 * step);               /* proposer.py:23 */
    } else {
        Jb.ca = sqrt(1 - step * step);        /* proposer.py:77 */
        Jb.cb = step;
    }
    Jb.prior_chol = prior_chol; Jb.u0 = u0; Jb.IC = IC; Jb.normals = normals; Jb.uniforms = uniforms;
    Jb.states = states; Jb.phi_v_out = phi_v_out; Jb.accepted_out = accepted_out; Jb.work = work;
    if (n_threads > n_chains) n_threads = n_chains;
    if (n_threads <= 1) {
        chain_worker(&Jb);
        return 0;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    int started = 0;
    for (int i = 0; i < n_threads; ++i)
        if (pthread_create(&th[i], NULL, chain_worker, &Jb) == 0) ++started; else break;
    if (started == 0) chain_worker(&Jb);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
    return 0;
}
