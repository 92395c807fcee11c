// Multi-warp ("team") variant of the Burgers solver for grids that do not fit one warp's
// registers: N = TM*32*CPL cells (2048 = 2x32x32, 4096 = 4x32x32), one CTA of TM warps per chain.
// Same arithmetic as BurgersWarp (burgers.cuh) -- EXACT stays bit-identical to the reference --
// plus, per time step, two CTA barriers and a shared-memory exchange of
//   * each warp's first / last cell (halo for the neighbouring warp; the flux at a warp boundary
//     is recomputed by the right-hand warp from the halo cell, no flux is exchanged), and
//   * each warp's CFL key (max |u| as an ordered integer).
// Buffer A (step start) and buffer B (between the SSPRK2 stages) alternate, so one barrier per
// exchange is enough: a buffer is rewritten only after every warp has passed the *other* barrier.
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
    __device__ __forceinline__ void flux_fused(const double (&w)[CPL], double wL, double wR, bool left_general,
                                               int lane, double (&F)[CPL], double &Fl) {
        double s[CPL + 1];
#pragma unroll
        for (int k = 0; k < CPL; ++k) s[k] = w[k] * w[k];
        double wr = shfl_down1(w[0]);
        wr = (lane == 31) ? wR : wr;
        s[CPL] = wr * wr;
        F[CPL - 1] = W1::flux2(w[CPL - 1], s[CPL - 1], wr, s[CPL]);
        Fl = shfl_up1(F[CPL - 1]);
#pragma unroll
        for (int k = 0; k < CPL - 1; ++k) F[k] = W1::flux2(w[k], s[k], w[k + 1], s[k + 1]);
        const double Fg = W1::flux2(wL, wL * wL, w[0], s[0]);
        const double Fb = left_general ? Fg : s[0];
        Fl = (lane == 0) ? Fb : Fl;
    }

    template <bool FIRST>
    __device__ __forceinline__ double step_fused(const BurgersConsts &C, TeamXch &X, int tw, int lane) {
        const uint64_t wk = warp_key();
        if (lane == 0) X.key[tw] = wk;
        double wL, wR;
        exchange(X, 0, tw, lane, u, FIRST ? gL : 0.0, FIRST ? gR : u[CPL - 1], wL, wR);
        uint64_t mk = X.key[0];
#pragma unroll
        for (int w = 1; w < TM; ++w) mk = key_max(mk, X.key[w]);
        const double m = __hiloint2double((int)(mk >> 32), (int)(uint32_t)mk);
        const double dt = C.half_dx * W1::fast_rcp(m);
        const double c8 = dt * C.c8_scale;
        double F[CPL], Fl, th[CPL], us[CPL];
        flux_fused(u, wL, wR, FIRST || tw > 0, lane, F, Fl);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const double dF = F[k] - (k == 0 ? Fl : F[k - 1]);
            th[k] = fma(c8, dF, u[k]);
            us[k] = fma(c8, dF, th[k]);
        }
        exchange(X, 1, tw, lane, us, 0.0, us[CPL - 1], wL, wR);
        flux_fused(us, wL, wR, tw > 0, lane, F, Fl);
#pragma unroll
        for (int k = 0; k < CPL; ++k) u[k] = fma(c8, F[k] - (k == 0 ? Fl : F[k - 1]), th[k]);
        return dt;
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

    template <bool FIRST, bool POW2>
    __device__ __forceinline__ double step(const BurgersConsts &C, TeamXch &X, int tw, int lane) {
        if (NUMERICS == NUM_FUSED) return step_fused<FIRST>(C, X, tw, lane);
        return step_exact<FIRST, POW2>(C, X, tw, lane);
    }

    template <bool POW2>
    __device__ __forceinline__ int time_loop(const BurgersConsts &C, TeamXch &X, int tw, int lane) {
        double t = 0.0;   // every warp of the team computes the same dt, hence the same trip count
        int n = 0;
        if (t < C.T && n < C.max_fv_steps) {
            t += step<true, POW2>(C, X, tw, lane);
            ++n;
        }
        while (t < C.T && n < C.max_fv_steps) {
            t += step<false, POW2>(C, X, tw, lane);
            ++n;
        }
        capped = t < C.T;
        return n;
    }

    __device__ __forceinline__ int integrate(const BurgersDev &B, TeamXch &X, double pi, int tw, int lane) {
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
        BurgersConsts C;
        C.T = B.T;
        C.half_dx = B.half_dx;
        C.neg_inv_dx = B.neg_inv_dx;
        C.neg_dx = -B.dx;
        C.c8_scale = 0.25 * B.neg_inv_dx;
        C.N = N;
        C.max_fv_steps = B.max_fv_steps;
        if (NUMERICS == NUM_FUSED || B.dx_pow2) return time_loop<true>(C, X, tw, lane);
        return time_loop<false>(C, X, tw, lane);
    }
};

}  // namespace ipmcmc
