// Multi-warp ("team") variant of the Burgers solver for grids that do not fit one warp's
// registers: N = TM*32*CPL cells (2048 = 2x32x32, 4096 = 4x32x32), one CTA of TM warps per chain.
// Same arithmetic as BurgersWarp (burgers.cuh) -- EXACT stays bit-identical to the reference --
// plus, per time step, two CTA barriers and a shared-memory exchange of
//   * each warp's first / last cell (halo for the neighbouring warp; the flux at a warp boundary
//     is recomputed by the right-hand warp from the halo cell, no flux is exchanged), and
//   * each warp's CFL key (max |u| as an ordered integer).
// Buffer A (step start) and buffer B (between the SSPRK2 stages) alternate, so one barrier per
// exchange is enough: a buffer is rewritten only after every warp has passed the *other* barrier.
// FUSED numerics take the two shortcuts of the one-warp solver (burgers.cuh), decided per solve on the state
// after the first time step and uniform over the CTA: positive states run the select-free flux, and monotone
// states take max|u| from the two end cells of the grid -- which the halo exchange has already put into
// shared memory, so the per-warp key tree, its two REDUX reductions and the key exchange disappear from the
// time step altogether (end-of-solve guard and fallback as in the one-warp solver).
#pragma once
#include "burgers.cuh"

namespace ipmcmc {

struct TeamXch {
    double first[2][4], last[2][4];
    unsigned long long key[4];
    double phi;
    int n_fv, capped;
};

template <int CPL, int NUMERICS, int TM>
struct BurgersTeam {
    using W1 = BurgersWarp<CPL, NUMERICS, false>;
    double u[CPL];
    double gL, gR;
    bool capped;

    __device__ __forceinline__ uint64_t warp_key() const {
        uint64_t key[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) key[k] = abs_key(u[k]);
#pragma unroll
        for (int w = 1; w < CPL; w *= 2)
#pragma unroll
            for (int k = 0; k + w < CPL; k += 2 * w) key[k] = key_max(key[k], key[k + w]);
        const uint32_t hi = (uint32_t)(key[0] >> 32), lo = (uint32_t)key[0];
        const uint32_t mh = __reduce_max_sync(FULL, hi);
        const uint32_t ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
        return ((uint64_t)mh << 32) | ml;
    }

    // publish first/last cell of w into buffer b, barrier, fetch the neighbours' halo cells
    __device__ __forceinline__ void exchange(TeamXch &X, int b, int tw, int lane, const double (&w)[CPL],
                                             double ghostL, double ghostR, double &wL, double &wR) {
        if (lane == 0) X.first[b][tw] = w[0];
        if (lane == 31) X.last[b][tw] = w[CPL - 1];
        __syncthreads();
        wL = (tw > 0) ? X.last[b][tw - 1] : ghostL;
        wR = (tw < TM - 1) ? X.first[b][tw + 1] : ghostR;
    }

    // ---- FUSED ------------------------------------------------------------------------------
    bool positive, monotone;   // CTA-uniform, decided after the first time step (decide())
    bool mono_ok;              // end-of-solve guard of the monotone shortcut (set at the end of fused_loop)

    template <bool POS>
    __device__ __forceinline__ void flux_fused(const double (&w)[CPL], double wL, double wR, bool left_general,
                                               int lane, double (&F)[CPL], double &Fl) {
        double s[CPL + 1];
#pragma unroll
        for (int k = 0; k < CPL; ++k) s[k] = w[k] * w[k];
        double wr = shfl_down1(w[0]);
        wr = (lane == 31) ? wR : wr;
        s[CPL] = wr * wr;
        if (POS) {   // every cell > 0: the upwind side is always the left cell (burgers.cuh, flux_fused<.., POS>)
            const double dl = wr - w[CPL - 1];
            F[CPL - 1] = fma(dl * fabs(dl), -0.5, s[CPL - 1]);
        } else {
            F[CPL - 1] = W1::flux2(w[CPL - 1], s[CPL - 1], wr, s[CPL]);
        }
        Fl = shfl_up1(F[CPL - 1]);
#pragma unroll
        for (int k = 0; k < CPL - 1; ++k) {
            if (POS) {
                const double dk = w[k + 1] - w[k];
                F[k] = fma(dk * fabs(dk), -0.5, s[k]);
            } else {
                F[k] = W1::flux2(w[k], s[k], w[k + 1], s[k + 1]);
            }
        }
        double Fg;
        if (POS) {
            const double d0 = w[0] - wL;
            Fg = fma(d0 * fabs(d0), -0.5, wL * wL);
        } else {
            Fg = W1::flux2(wL, wL * wL, w[0], s[0]);
        }
        const double Fb = left_general ? Fg : s[0];
        Fl = (lane == 0) ? Fb : Fl;
    }

    template <bool FIRST, bool POS, bool MONO>
    __device__ __forceinline__ double step_fused(const BurgersConsts &C, TeamXch &X, int tw, int lane) {
        if (!MONO) {
            const uint64_t wk = warp_key();
            if (lane == 0) X.key[tw] = wk;
        }
        double wL, wR;
        exchange(X, 0, tw, lane, u, FIRST ? gL : 0.0, FIRST ? gR : u[CPL - 1], wL, wR);
        double m;
        if (MONO) {   // monotone profile: |u| is largest at one of the two ends of the grid
            m = absmax_bits(X.first[0][0], X.last[0][TM - 1]);
        } else {
            uint64_t mk = X.key[0];
#pragma unroll
            for (int w = 1; w < TM; ++w) mk = key_max(mk, X.key[w]);
            m = __hiloint2double((int)(mk >> 32), (int)(uint32_t)mk);
        }
        // dt = half_dx / m and c8 = dt * c8_scale from one cubic correction of the MUFU seed (BurgersWarp::fused_dt)
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(m));
        const double e = fma(-m, r, 1.0);
        const double rd = r * C.half_dx, rc = r * C.k8;
        const double tt = fma(e, e, e);
        const double dt = fma(rd, tt, rd), c8 = fma(rc, tt, rc);
        double F[CPL], Fl, th[CPL], us[CPL];
        flux_fused<POS && !FIRST>(u, wL, wR, FIRST || tw > 0, lane, F, Fl);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const double dF = F[k] - (k == 0 ? Fl : F[k - 1]);
            th[k] = fma(c8, dF, u[k]);
            us[k] = fma(c8, dF, th[k]);
        }
        exchange(X, 1, tw, lane, us, 0.0, us[CPL - 1], wL, wR);
        flux_fused<POS>(us, wL, wR, tw > 0, lane, F, Fl);
#pragma unroll
        for (int k = 0; k < CPL; ++k) u[k] = fma(c8, F[k] - (k == 0 ? Fl : F[k - 1]), th[k]);
        return dt;
    }

    // positive / monotone over the WHOLE grid (all warps of the team), on the state after the first time step
    __device__ __forceinline__ void decide(TeamXch &X, int tw, int lane, bool allow_mono) {
        double wL, wR;
        exchange(X, 0, tw, lane, u, u[0], u[CPL - 1], wL, wR);   // wR: first cell of the next warp
        double wr = shfl_down1(u[0]);
        wr = (lane == 31) ? wR : wr;
        bool pos = true, ni = u[CPL - 1] >= wr, nd = u[CPL - 1] <= wr;
#pragma unroll
        for (int k = 0; k < CPL; ++k) pos = pos && (u[k] > 0.0);
#pragma unroll
        for (int k = 0; k + 1 < CPL; ++k) {
            ni = ni && (u[k] >= u[k + 1]);
            nd = nd && (u[k] <= u[k + 1]);
        }
        positive = __syncthreads_and(pos) != 0;
        const bool ni_all = __syncthreads_and(ni) != 0, nd_all = __syncthreads_and(nd) != 0;
        monotone = IPMCMC_MONO && allow_mono && (ni_all || nd_all);
    }
    // end-of-solve guard of the monotone shortcut (BurgersWarp::mono_end_ok)
    __device__ __forceinline__ bool mono_end_ok(TeamXch &X, int tw, int lane) {
        const uint64_t wk = warp_key();
        if (lane == 0) {
            X.key[tw] = wk;
            X.first[0][tw] = u[0];
        }
        if (lane == 31) X.last[0][tw] = u[CPL - 1];
        __syncthreads();
        uint64_t mk = X.key[0];
#pragma unroll
        for (int w = 1; w < TM; ++w) mk = key_max(mk, X.key[w]);
        const double m_all = __hiloint2double((int)(mk >> 32), (int)(uint32_t)mk);
        const double m_end = absmax_bits(X.first[0][0], X.last[0][TM - 1]);
        __syncthreads();
        return m_all <= m_end * (1.0 + 1e-12);
    }

    // ---- EXACT ------------------------------------------------------------------------------
    template <bool SECOND, bool POW2>
    __device__ __forceinline__ void stage_exact(const BurgersConsts &C, const double (&w)[CPL], double wL, double wR,
                                                bool left_general, double dt, int lane, const double (&aux)[CPL],
                                                double (&out)[CPL]) {
        double h[CPL + 1], g[CPL + 1], F[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            h[k] = 0.5 * w[k];
            g[k] = h[k] * h[k];
        }
        double wr = shfl_down1(w[0]);
        wr = (lane == 31) ? wR : wr;
        h[CPL] = 0.5 * wr;
        g[CPL] = h[CPL] * h[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const double ur = (k + 1 < CPL) ? w[k + 1] : wr;
            F[k] = W1::flux_exact(w[k], h[k], g[k], ur, h[k + 1], g[k + 1]);
        }
        double Fl = shfl_up1(F[CPL - 1]);
        const double hl = 0.5 * wL;
        const double Fg = W1::flux_exact(wL, hl, hl * hl, w[0], h[0], g[0]);
        const double Fb = left_general ? Fg : g[0] + g[0];
        Fl = (lane == 0) ? Fb : Fl;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const double dF = F[k] - (k == 0 ? Fl : F[k - 1]);
            const double dudt = POW2 ? dF * C.neg_inv_dx : dF / C.neg_dx;
            const double inc = dt * dudt;
            if (!SECOND) {
                out[k] = w[k] + inc;
            } else {
                const double ustar = w[k] + inc;
                out[k] = (aux[k] + ustar) * 0.5;
            }
        }
    }

    template <bool FIRST, bool POW2>
    __device__ __forceinline__ double step_exact(const BurgersConsts &C, TeamXch &X, int tw, int lane) {
        const uint64_t wk = warp_key();
        if (lane == 0) X.key[tw] = wk;
        double wL, wR;
        exchange(X, 0, tw, lane, u, FIRST ? gL : 0.0, FIRST ? gR : u[CPL - 1], wL, wR);
        uint64_t mk = X.key[0];
#pragma unroll
        for (int w = 1; w < TM; ++w) mk = key_max(mk, X.key[w]);
        const double m = __hiloint2double((int)(mk >> 32), (int)(uint32_t)mk);
        const double dt = C.half_dx / m;
        double us[CPL], un[CPL];
        stage_exact<false, POW2>(C, u, wL, wR, FIRST || tw > 0, dt, lane, u, us);
        exchange(X, 1, tw, lane, us, 0.0, us[CPL - 1], wL, wR);
        stage_exact<true, POW2>(C, us, wL, wR, tw > 0, dt, lane, u, un);
#pragma unroll
        for (int k = 0; k < CPL; ++k) u[k] = un[k];
        return dt;
    }

    template <bool POS, bool MONO>
    __device__ __forceinline__ int fused_loop(const BurgersConsts &C, TeamXch &X, int tw, int lane, double t, int n) {
        while (t < C.T && n < C.max_fv_steps) {
            t += step_fused<false, POS, MONO>(C, X, tw, lane);
            ++n;
        }
        capped = t < C.T;
        // the guard of the monotone shortcut is evaluated HERE, so that `monotone` / `positive` are dead behind the
        // dispatch and hold no predicate register across the time loop (burgers.cuh, time_loop_mono)
        mono_ok = !MONO || capped || mono_end_ok(X, tw, lane);
        return n;
    }

    template <bool POW2>
    __device__ __forceinline__ int time_loop(const BurgersConsts &C, TeamXch &X, int tw, int lane, bool allow_mono) {
        double t = 0.0;   // every warp of the team computes the same dt, hence the same trip count
        int n = 0;
        positive = monotone = false;
        mono_ok = true;
        if (NUMERICS == NUM_FUSED) {
            if (t < C.T && n < C.max_fv_steps) {
                t += step_fused<true, false, false>(C, X, tw, lane);
                ++n;
                decide(X, tw, lane, allow_mono);
            }
            if (positive) {
                if (monotone) return fused_loop<true, true>(C, X, tw, lane, t, n);
                return fused_loop<true, false>(C, X, tw, lane, t, n);
            }
            if (monotone) return fused_loop<false, true>(C, X, tw, lane, t, n);
            return fused_loop<false, false>(C, X, tw, lane, t, n);
        }
        if (t < C.T && n < C.max_fv_steps) {
            t += step_exact<true, POW2>(C, X, tw, lane);
            ++n;
        }
        while (t < C.T && n < C.max_fv_steps) {
            t += step_exact<false, POW2>(C, X, tw, lane);
            ++n;
        }
        capped = t < C.T;
        return n;
    }

    __device__ __forceinline__ void init_state(const BurgersDev &B, double pi, int tw, int lane) {
        const int N = B.N;
        const double p_left = shfl(pi, 0), p_right = shfl(pi, 1), p_jump = shfl(pi, 2);
        const double left = 1.0 + p_left;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int c = (tw * 32 + lane) * CPL + k;
            u[k] = (B.x[c + 1] < p_jump) ? left : p_right;
        }
        gL = (B.x[0] < p_jump) ? left : p_right;
        gR = (B.x[N + 1] < p_jump) ? left : p_right;
        for (int m = 0; m < B.n_modes; ++m) {
            const double a = shfl(pi, 3 + m);
            const double *phi = B.basis + (size_t)m * (N + 2);
#pragma unroll
            for (int k = 0; k < CPL; ++k) u[k] = u[k] + a * phi[(tw * 32 + lane) * CPL + k + 1];
            gL = gL + a * phi[0];
            gR = gR + a * phi[N + 1];
        }
    }

    __device__ __forceinline__ int integrate(const BurgersDev &B, TeamXch &X, double pi, int tw, int lane) {
        BurgersConsts C;
        C.T = C.T_reg = B.T;
        C.half_dx = B.half_dx;
        C.neg_inv_dx = B.neg_inv_dx;
        C.neg_dx = -B.dx;
        C.c8_scale = 0.25 * B.neg_inv_dx;
        C.k8 = C.half_dx * C.c8_scale;
        C.N = B.N;
        C.max_fv_steps = B.max_fv_steps;
        int n = 0;
        for (int pass = 0; pass < 2; ++pass) {   // second pass only if the guard of the monotone shortcut fails
            init_state(B, pi, tw, lane);
            const bool allow_mono = pass == 0 && !B.no_mono;
            if (NUMERICS == NUM_FUSED || B.dx_pow2) n = time_loop<true>(C, X, tw, lane, allow_mono);
            else n = time_loop<false>(C, X, tw, lane, allow_mono);
            if (mono_ok) break;
        }
        return n;
    }
};

}  // namespace ipmcmc
