// Two-scale Lorenz-96 + Dormand-Prince 5(4) with scipy's controller, laid out for the warp:
// ONE LANE PER SLOW VARIABLE.  Lane k of a chain's lane group holds X_k and its J fast variables
// Y_{k,0..J-1} in registers, so the fast-fast coupling (periodic WITHIN a block, lorenz.py:90-101),
// the block mean and the slow->fast forcing are thread-local; only the three slow neighbours
// X_{k-1}, X_{k-2}, X_{k+1} travel by shuffle (6 SHFL.32 per RHS instead of 14+ for a
// lane-per-variable layout).  A chain occupies K lanes; floor(32/K) chains share a warp (5 chains
// for the reference's K = 6) and step in lock-step through predicated attempts.
//
// Restates Lorenz96.__call__ (report/scripts/lorenz.py:44-101), moment_function
// (report/scripts/lorenz_mcmc.py:17-40), LorenzObservationOperator (lorenz_mcmc.py:43-71) and
// scipy.integrate.solve_ivp(method='RK45') with default tolerances (scipy/integrate/_ivp/rk.py:14-175,
// common.py:63-140; call site lorenz_mcmc.py:70-71).  CPU statement: oracle/lorenz_np.py.
#pragma once
#include "common.cuh"

// developer switch for A/B measurements (tools/build_lorenz_variants.py): the two-level group sum of the FUSED path
// max(|y|,|y_new|) of the FUSED error scale as DSETP on |.| operands + one 64-bit select (1) or as integer keys (0)
#ifndef IPMCMC_LORENZ_ABSMAX_DSETP
#define IPMCMC_LORENZ_ABSMAX_DSETP 1
#endif
#ifndef IPMCMC_LORENZ_GSUM2
#define IPMCMC_LORENZ_GSUM2 1
#endif

namespace ipmcmc {

struct LorenzDev {
    int K, J, max_attempts, nvar;
    double T, c, rtol, atol;
    double param_mean[3];
    PotentialDev pot;
};

// Dormand-Prince tableau (scipy rk.py RK45.A/B/C/E) in constant memory: the FMAs take the
// coefficient as a constant-bank operand (no UMOV pairs to materialise 64-bit immediates, and a
// two-register DFMA issues every 2 cycles instead of 3).
enum : int { A21, A31, A32, A41, A42, A43, A51, A52, A53, A54, A61, A62, A63, A64, A65,
             B1, B3, B4, B5, B6, E1, E3, E4, E5, E6, E7, DP_N };
__constant__ double DP[DP_N] = {
    1.0 / 5, 3.0 / 40, 9.0 / 40, 44.0 / 45, -56.0 / 15, 32.0 / 9,
    19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729,
    9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656,
    35.0 / 384, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84,
    -71.0 / 57600, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
#define RK_SAFETY 0.9
#define RK_MIN_FACTOR 0.2
#define RK_MAX_FACTOR 10.0

// 1/x to ~1 ulp: MUFU.RCP64H seed + two Newton rounds (branch-free, unlike the IEEE division)
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

// Parameters of one chain's ODE, identical on all lanes of its group.
struct LorenzTheta {
    double F, h, c, b;
    double hc, hJ;  // h*c (lorenz.py:42) and h/J (lorenz.py:98), hoisted out of the RHS
    // FUSED numerics integrate the SCALED fast variables Z = s*Y, s = -b*c (rhs_fused): s, 1/s,
    // s*c*h/J (forcing of Z by X) and -(h*c/J)/s (coupling of X to sum Z)
    double s, inv_s, sA, mS;
    __device__ __forceinline__ void finish(int J) {
        hc = h * c;
        hJ = h / (double)J;
        // |s| is kept away from zero: below 1e-30 the quadratic term s*Y*dY is far under the rounding of c*Y
        // whatever the scale, so the clamped scale integrates the same equations to rounding (NaN stays NaN)
        const double s0 = -(b * c);
        s = (fabs(s0) < 1e-30) ? copysign(1e-30, s0) : s0;
        inv_s = 1.0 / s;
        sA = s * (c * hJ);
        mS = -(hc / (double)J) * inv_s;
    }
};

// x^(-1/10) for x in [1e-30, 1e30]: fp32 seed on the SFU (MUFU.LG2 / MUFU.EX2, relative error ~1e-6)
// and ROUNDS Newton rounds y <- y + y*(1 - x*y^10)/10 (error -> 5.5 e^2 each): two give ~1 ulp (14 fp64
// instructions in a chain of 12, instead of the ~70-instruction log + exp of pow()); one gives ~5e-12,
// plenty for a step-size factor and 50 cycles shorter on the attempt's critical path (FUSED numerics).
template <int ROUNDS>
__device__ __forceinline__ double pow_m01(double x) {
    double y = (double)exp2f(-0.1f * __log2f((float)x));
#pragma unroll
    for (int it = 0; it < ROUNDS; ++it) {
        const double y2 = y * y, y4 = y2 * y2, y8 = y4 * y4, y10 = y8 * y2;
        const double r = fma(-x, y10, 1.0);
        y = fma(y * 0.1, r, y);
    }
    return y;
}

// FUSED controller: h * 0.9 * x^(-1/10) for a positive normal x inside the fp32 range, ~5e-13 relative.
// Seed on the SFU in fp32 WITHOUT the slow F2F conversions or the denormal fix-ups of __log2f/exp2f
// (x is re-biased into an fp32 bit pattern with two integer instructions, the result widened the same way;
// lg2/ex2.approx.ftz), then one Newton round y <- y (1 + (1 - x y^10)/10) arranged in 5 dependent levels
// (y2 | y4, x*y2 | y8 | 1 - (x y2) y8 | fma) with SAFETY and the step h folded into the last FMA (h*0.9y and
// h*0.09y form beside the chain), so that the next step size is ONE instruction behind the residual.
// Garbage in (0, inf, NaN, out of the fp32 range) gives garbage out: the caller overrides those cases by
// predicates on x.
__device__ __forceinline__ double rk_hfactor_fused(double ss, double inv_n, float log2n_10, double h) {
    // seed straight from the bits of ss = n*e2: (ss/n)^(-1/10) = 2^(-0.1 lg2(ss) + 0.1 lg2(n)), the 1/n as one FFMA
    const uint32_t hi = (uint32_t)__double2hiint(ss), lo = (uint32_t)__double2loint(ss);
    const float xf = __uint_as_float(((hi - 0x38000000u) << 3) | (lo >> 29));      // truncated, 2^-23 relative
    float lg, yf;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(xf));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(fmaf(lg, -0.1f, log2n_10)));
    const uint32_t yb = __float_as_uint(yf);
    const double y = __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));   // exact widening
    const double x = ss * inv_n;                                                   // beside the SFU chain
    const double h09 = h * RK_SAFETY, h009 = h * (0.1 * RK_SAFETY);                // beside the SFU chain
    const double y2 = y * y, hy09 = y * h09, hy009 = y * h009;
    const double y4 = y2 * y2, a = x * y2;
    const double y8 = y4 * y4;
    const double r = fma(-a, y8, 1.0);
    return fma(hy009, r, hy09);
}

// SAFETY * error_norm ** (-1/5) (scipy rk.py:157,166) from the SQUARED norm e2 = error_norm^2.
// Outside [1e-30, 1e30] the caller's clamps (MIN_FACTOR 0.2, MAX_FACTOR 10) decide anyway, so the
// argument is clamped first; NaN stays NaN (a NaN error norm rejects the step and shrinks h by
// MIN_FACTOR, as in scipy: nan < 1 is False, max(0.2, nan) is 0.2).
template <int ROUNDS>
__device__ __forceinline__ double rk_raw_factor(double e2) {
    const double x = fmin(fmax(e2, 1e-30), 1e30);
    const double f = RK_SAFETY * pow_m01<ROUNDS>(x);
    return (e2 != e2) ? e2 : f;
}

// KT: K known at compile time (6 = the reference's problem) or 0 = run-time K.
// NUM: NUM_EXACT (reference rounding order in the RHS) or NUM_FUSED (contracted RHS).
enum : int { LNUM_EXACT = 0, LNUM_FUSED = 1 };

template <int J, int KT = 0, int NUM = LNUM_EXACT>
struct LorenzLanes {
    static constexpr int NV = J + 1;  // variables per lane: y[0] = X_k, y[1+j] = Y_{k,j}
    int src_m1, src_m2, src_p1;       // absolute lane ids of X_{k-1}, X_{k-2}, X_{k+1}
    int base, K, k;
    bool valid;

    __device__ __forceinline__ void init(int lane, int K_, int groups) {
        K = K_;
        const int g = lane / K_;
        valid = g < groups;
        base = valid ? g * K_ : 0;
        k = valid ? lane - base : 0;
        src_m1 = base + (k + K_ - 1) % K_;
        src_m2 = base + (k + 2 * K_ - 2) % K_;
        src_p1 = base + (k + 1) % K_;
    }

    // sum over the K lanes of the group: shuffle-down tree into the group's first lane, then one
    // broadcast, so every lane of the chain holds the SAME bits (control flow stays group-uniform)
    // K known at compile time: every lane gathers the K partial sums with K independent
    // shuffles and adds them as a balanced tree in lane order (identical bits on every lane, ~3
    // dependent DADDs) instead of log2(K) dependent shuffle+add rounds plus a broadcast.
    __device__ __forceinline__ double group_sum(double v) const {
        if (KT > 0) {
            constexpr int KG = KT > 0 ? KT : 1;
            double g[KG];
#pragma unroll
            for (int i = 0; i < KG; ++i) g[i] = __shfl_sync(FULL, v, base + i);
#pragma unroll
            for (int w = 1; w < KG; w *= 2)
#pragma unroll
                for (int i = 0; i + w < KG; i += 2 * w) g[i] = g[i] + g[i + w];
            return g[0];
        }
        double s = v;
        for (int off = 1; off < K; off *= 2) {
            const double t = __shfl_down_sync(FULL, s, off);
            if (k + off < K) s = s + t;
        }
        return __shfl_sync(FULL, s, base);
    }

    // FUSED numerics, even K known at compile time: neighbours add up first (a + b and b + a are the same bits,
    // so both lanes of a pair hold one value), then every lane gathers the K/2 pair sums and adds them in lane
    // order: 1 + K/2 shuffles of 64 bits instead of K (K = 6: 8 SHFL.32 instead of 12, and the gather no
    // longer queues six deep in front of the step-size controller).  Identical bits on every lane of the group.
    __device__ __forceinline__ double group_sum_fast(double v) const {
        if (IPMCMC_LORENZ_GSUM2 && KT > 0 && KT % 2 == 0) {
            constexpr int KH = KT > 0 ? KT / 2 : 1;
            const double p = v + __shfl_xor_sync(FULL, v, 1);      // base is even: the partner is in the group
            double g[KH];
#pragma unroll
            for (int i = 0; i < KH; ++i) g[i] = __shfl_sync(FULL, p, base + 2 * i);
#pragma unroll
            for (int w = 1; w < KH; w *= 2)
#pragma unroll
                for (int i = 0; i + w < KH; i += 2 * w) g[i] = g[i] + g[i + w];
            return g[0];
        }
        return group_sum(v);
    }

    // FUSED numerics integrate Z_j = s*Y_j (s = -b*c, LorenzTheta): state in / out of the scaled variables
    __device__ __forceinline__ void to_scaled(const LorenzTheta &th, double (&y)[NV]) const {
        if (NUM == LNUM_FUSED) {
#pragma unroll
            for (int j = 1; j < NV; ++j) y[j] = y[j] * th.s;
        }
    }
    __device__ __forceinline__ void from_scaled(const LorenzTheta &th, double (&y)[NV]) const {
        if (NUM == LNUM_FUSED) {
#pragma unroll
            for (int j = 1; j < NV; ++j) y[j] = y[j] * th.inv_s;
        }
    }

    // FUSED numerics: the same right-hand side contracted for the fp64 pipe, in the scaled fast variables
    // Z_j = s*Y_j, s = -b*c, which turn the quadratic coefficient into 1 (20 instead of 45 instructions at J = 4;
    // 24 with the products s*Y_{j+1} formed per evaluation):
    //   dY_j = c*(-Y_j - b*Y_{j+1}*(Y_{j+2} - Y_{j-1}) + (h/J) X)
    //   dZ_j = s*dY_j = Z_{j+1}*(Z_{j+2} - Z_{j-1}) - c*Z_j + (s*c*h/J) X  = fma(Z_{j+1}, Z_{j+2}-Z_{j-1}, fma(-c, Z_j, sA*X))
    //   dX   = F - X - X_{k-1}*(X_{k-2} - X_{k+1}) - (hc/J) sum_j Y_j,   (hc/J) sum Y = -mS * sum Z
    // y[0] = X_k, y[1+j] = Z_{k,j}; dy likewise (d/dt of the scaled variables).
    // C1: the time-scale ratio c is the reference's 1 (lorenz_mcmc.py:91, theta = [F, h, c, b] = [10, 10, 1, 10]):
    // -c*Z_j + A is the subtraction A - Z_j (same bits as the FMA with c = 1; two register operands, a two-cycle
    // issue) instead of an FMA with c in a register that the compiler cannot keep uniform beside the tableau.
    template <bool C1 = false>
    __device__ __forceinline__ void rhs_fused(const LorenzTheta &th, const double (&y)[NV], double (&dy)[NV]) const {
        const double X = y[0];
        const double Xm1 = __shfl_sync(FULL, X, src_m1);
        const double Xm2 = __shfl_sync(FULL, X, src_m2);
        const double Xp1 = __shfl_sync(FULL, X, src_p1);
        double out = th.F - X;
        if (J > 0) {
            const double A = th.sA * X;
            double t[J];
#pragma unroll
            for (int j = 0; j < J; ++j) t[j] = y[1 + j];
#pragma unroll
            for (int w = 1; w < J; w *= 2)
#pragma unroll
                for (int j = 0; j + w < J; j += 2 * w) t[j] = t[j] + t[j + w];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double e = y[1 + (j + 2) % J] - y[1 + (j + J - 1) % J];
                dy[1 + j] = fma(y[1 + (j + 1) % J], e, C1 ? (A - y[1 + j]) : fma(-th.c, y[1 + j], A));
            }
            out = fma(th.mS, t[0], out);
        }
        dy[0] = fma(-Xm1, Xm2 - Xp1, out);
    }

    // d(state)/dt in the reference's rounding order (lorenz.py:73-101)
    __device__ __forceinline__ void rhs(const LorenzTheta &th, const double (&y)[NV], double (&dy)[NV]) const {
        if (NUM == LNUM_FUSED) return rhs_fused(th, y, dy);
        const double X = y[0];
        const double Xm1 = __shfl_sync(FULL, X, src_m1);
        const double Xm2 = __shfl_sync(FULL, X, src_m2);
        const double Xp1 = __shfl_sync(FULL, X, src_p1);
        double out = -X;
        out = out - (Xm1 * Xm2 - Xm1 * Xp1);
        out = out + th.F;
        if (J > 0) {
            // np.average(Y_k): numpy pairwise sum (sequential for J < 8) / J
            double s;
            if (J < 8) {
                s = -0.0;
#pragma unroll
                for (int j = 0; j < J; ++j) s = s + y[1 + j];
            } else {
                s = np_pairwise_sum([&](int j) { return y[1 + j]; }, 0, J);
            }
            out = out - th.hc * (s / (double)J);
            const double hx = th.hJ * X;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double Yp1 = y[1 + (j + 1) % J], Yp2 = y[1 + (j + 2) % J], Ym1 = y[1 + (j + J - 1) % J];
                double o = -y[1 + j];
                o = o - th.b * (Yp1 * Yp2 - Ym1 * Yp1);
                o = o + hx;
                dy[1 + j] = o * th.c;
            }
        }
        dy[0] = out;
    }

    // RMS norm over the chain's n = K*(J+1) variables of v/scale (scipy common.py:63-65)
    __device__ __forceinline__ double sumsq(const double (&v)[NV]) const {
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) ss = fma(v[i], v[i], ss);
        return group_sum(ss);
    }
    __device__ __forceinline__ double rms(const double (&v)[NV], double inv_sqrt_n) const {
        return sqrt(sumsq(v)) * inv_sqrt_n;
    }

    // One Dormand-Prince attempt (scipy rk.py:14-72 + error norm :111-116).
    // k1 = f(y) on entry; k7 = f(y_new) on exit.  Stage sums use FMA chains.
    // Returns the SUM OF SQUARES of error/scale over the chain's variables (error_norm^2 * n).
    __device__ __forceinline__ double attempt(const LorenzTheta &th, const double (&y)[NV], const double (&k1)[NV],
                                              double h, double rtol, double atol,
                                              double (&ynew)[NV], double (&k7)[NV]) const {
        double k2[NV], k3[NV], k4[NV], k5[NV], k6[NV], ys[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) ys[i] = fma(DP[A21] * k1[i], h, y[i]);
        rhs(th, ys, k2);
#pragma unroll
        for (int i = 0; i < NV; ++i) ys[i] = fma(fma(DP[A32], k2[i], DP[A31] * k1[i]), h, y[i]);
        rhs(th, ys, k3);
#pragma unroll
        for (int i = 0; i < NV; ++i) ys[i] = fma(fma(DP[A43], k3[i], fma(DP[A42], k2[i], DP[A41] * k1[i])), h, y[i]);
        rhs(th, ys, k4);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            ys[i] = fma(fma(DP[A54], k4[i], fma(DP[A53], k3[i], fma(DP[A52], k2[i], DP[A51] * k1[i]))), h, y[i]);
        rhs(th, ys, k5);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            ys[i] = fma(fma(DP[A65], k5[i],
                            fma(DP[A64], k4[i], fma(DP[A63], k3[i], fma(DP[A62], k2[i], DP[A61] * k1[i])))),
                        h, y[i]);
        rhs(th, ys, k6);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            ynew[i] = fma(h, fma(DP[B6], k6[i], fma(DP[B5], k5[i], fma(DP[B4], k4[i], fma(DP[B3], k3[i], DP[B1] * k1[i])))),
                          y[i]);
        rhs(th, ynew, k7);
        double e[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double err = fma(DP[E7], k7[i],
                                   fma(DP[E6], k6[i],
                                       fma(DP[E5], k5[i], fma(DP[E4], k4[i], fma(DP[E3], k3[i], DP[E1] * k1[i]))))) * h;
            const double scale = fma(absmax_bits(y[i], ynew[i]), rtol, atol);
            e[i] = err * rcp_newton(scale);
        }
        return sumsq(e);
    }

    // FUSED numerics: the same attempt arranged for the tail of the dependent chain.  Everything of the error
    // estimate that does not need k7 = f(y_new) -- the partial sum of the first six stages, the scale, its
    // reciprocal (MUFU seed + ONE Newton round: 2^-46) times h -- is formed while the last right-hand side is
    // in flight, so that ONE FMA per variable follows k7; the squares are summed as two chains.
    // Returns sum over the chain's variables of (error/scale)^2  (= n * error_norm^2).
    // `atol_z` = atol*|s|: the absolute tolerance of the scaled fast variables (error/scale is scale-free).
    // `ys2` = y + h*a21*k1, the state of the second stage (stage2()): the caller forms it at the END of the previous
    // attempt (rotated loop, attempts_fused), so that everything between the step-size controller and the first
    // right-hand side of the next attempt lies in ONE basic block and ptxas schedules the accept selects into the
    // shadow of the power instead of behind it.
    __device__ __forceinline__ void stage2(const double (&y)[NV], const double (&k1)[NV], double h, double (&ys2)[NV]) const {
#pragma unroll
        for (int i = 0; i < NV; ++i) ys2[i] = fma(DP[A21] * k1[i], h, y[i]);
    }
    template <bool C1 = false>
    __device__ __forceinline__ double attempt_fused(const LorenzTheta &th, const double (&y)[NV], const double (&k1)[NV],
                                                    const double (&ys2)[NV], double h, double rtol, double atol,
                                                    double atol_z, double (&ynew)[NV], double (&k7)[NV]) const {
        double k2[NV], k3[NV], k4[NV], k5[NV], k6[NV], ys[NV];
        rhs_fused<C1>(th, ys2, k2);
#pragma unroll
        for (int i = 0; i < NV; ++i) ys[i] = fma(fma(DP[A32], k2[i], DP[A31] * k1[i]), h, y[i]);
        rhs_fused<C1>(th, ys, k3);
#pragma unroll
        for (int i = 0; i < NV; ++i) ys[i] = fma(fma(DP[A43], k3[i], fma(DP[A42], k2[i], DP[A41] * k1[i])), h, y[i]);
        rhs_fused<C1>(th, ys, k4);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            ys[i] = fma(fma(DP[A54], k4[i], fma(DP[A53], k3[i], fma(DP[A52], k2[i], DP[A51] * k1[i]))), h, y[i]);
        rhs_fused<C1>(th, ys, k5);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            ys[i] = fma(fma(DP[A65], k5[i],
                            fma(DP[A64], k4[i], fma(DP[A63], k3[i], fma(DP[A62], k2[i], DP[A61] * k1[i])))),
                        h, y[i]);
        rhs_fused<C1>(th, ys, k6);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            ynew[i] = fma(h, fma(DP[B6], k6[i], fma(DP[B5], k5[i], fma(DP[B4], k4[i], fma(DP[B3], k3[i], DP[B1] * k1[i])))),
                          y[i]);
        rhs_fused<C1>(th, ynew, k7);
        double S[NV], Q[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double part = fma(DP[E6], k6[i], fma(DP[E5], k5[i], fma(DP[E4], k4[i], fma(DP[E3], k3[i], DP[E1] * k1[i]))));
#if IPMCMC_LORENZ_ABSMAX_DSETP
            // 15 instead of 36 instructions for the five maxima (a NaN y_new is selected and poisons the norm, as it must)
            const double ym = (fabs(y[i]) > fabs(ynew[i])) ? y[i] : ynew[i];
            const double scale = fma(fabs(ym), rtol, i == 0 ? atol : atol_z);
#else
            const double scale = fma(absmax_bits(y[i], ynew[i]), rtol, i == 0 ? atol : atol_z);
#endif
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(scale));
            r = fma(r, fma(-scale, r, 1.0), r);
            const double hr = h * r;
            S[i] = part * hr;
            Q[i] = DP[E7] * hr;
        }
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double e = fma(k7[i], Q[i], S[i]);
            if (i & 1) sb = fma(e, e, sb);
            else sa = fma(e, e, sa);
        }
        return group_sum_fast(sa + sb);
    }
};

// Per-chain integrator state for the lock-step solve.
template <int J, int KT = 0, int NUM = LNUM_EXACT>
struct LorenzSolve {
    static constexpr int NV = J + 1;
    double y[NV], f[NV];
    double msum[5];  // running sums of [X, Y0, X^2, X*Y0, Y0^2] over stored states (lorenz_mcmc.py:17-40)
    double t, h_abs;
    int n_t, n_acc, n_rej;
    bool done, step_rejected, new_step;

    __device__ __forceinline__ void add_moments() {
        const double X = y[0], Y0 = (J > 0) ? y[1] : 0.0;
        msum[0] += X;
        msum[1] += Y0;
        msum[2] += X * X;
        msum[3] += X * Y0;
        msum[4] += Y0 * Y0;
        ++n_t;
    }

    // solve_ivp(fun, (0,T), y0, 'RK45'): initial step (common.py:68-140), then steps until t == T.
    // `active`: this lane's chain takes part (others run predicated-off). All 32 lanes must call.
    // FUSED numerics work in the scaled fast variables Z = s*Y between the first and the last line (state,
    // derivatives and the three moment sums that contain Y_0); ratios to the error scale do not change because the
    // absolute tolerance of a scaled variable is scaled with it.
    __device__ __forceinline__ void solve(const LorenzLanes<J, KT, NUM> &L, const LorenzDev &P, const LorenzTheta &th,
                                          bool active) {
        const double inv_sqrt_n = 1.0 / sqrt((double)P.nvar);
        const double atol_z = (NUM == LNUM_FUSED) ? P.atol * fabs(th.s) : P.atol;
        t = 0.0;
        n_t = n_acc = n_rej = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) msum[i] = 0.0;
        L.to_scaled(th, y);
        L.rhs(th, y, f);
        add_moments();  // t0 is stored too (lorenz_mcmc.py:68)
        {   // select_initial_step, direction = +1, order = 4
            double sc[NV], a0[NV], a1[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                sc[i] = fma(fabs(y[i]), P.rtol, i == 0 ? P.atol : atol_z);
                a0[i] = y[i] / sc[i];
                a1[i] = f[i] / sc[i];
            }
            const double d0 = L.rms(a0, inv_sqrt_n), d1 = L.rms(a1, inv_sqrt_n);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = fmin(h0, P.T);
            double y1[NV], f1[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) y1[i] = fma(h0, f[i], y[i]);
            L.rhs(th, y1, f1);
#pragma unroll
            for (int i = 0; i < NV; ++i) a0[i] = (f1[i] - f[i]) / sc[i];
            const double d2 = L.rms(a0, inv_sqrt_n) / h0;
            double h1;
            if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
            else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
            h_abs = fmin(fmin(100.0 * h0, h1), P.T);
        }
        done = !active || !(P.T > 0.0);
        step_rejected = false;
        new_step = true;
        if (NUM == LNUM_FUSED) {
            if (P.c == 1.0) attempts_fused<true>(L, P, th, atol_z);   // kernel parameter: warp-uniform
            else attempts_fused<false>(L, P, th, atol_z);
            n_t = n_acc + 1;
            L.from_scaled(th, y);
            msum[1] *= th.inv_s;                 // sum Y_0
            msum[3] *= th.inv_s;                 // sum X Y_0
            msum[4] *= th.inv_s * th.inv_s;      // sum Y_0^2
        } else {
            attempts_exact(L, P, th);
        }
    }

    // ---- RungeKutta._step_impl (rk.py:118-175), one attempt per iteration, scipy's order of operations
    __device__ __forceinline__ void attempts_exact(const LorenzLanes<J, KT, NUM> &L, const LorenzDev &P, const LorenzTheta &th) {
        const double inv_n = 1.0 / (double)P.nvar;
        int guard = 0;
        while (__any_sync(FULL, !done)) {
            const double t_next = __longlong_as_double(__double_as_longlong(t) + 1);  // nextafter(t, +inf), t >= 0
            const double min_step = 10.0 * fabs(t_next - t);
            if (new_step) {
                if (h_abs < min_step) h_abs = min_step;
                step_rejected = false;
            }
            bool fail = h_abs < min_step;  // TOO_SMALL_STEP: solve_ivp stops with what it has
            double h = h_abs;
            double t_new = t + h;
            if (t_new - P.T > 0.0) t_new = P.T;
            h = t_new - t;
            const double h_abs_used = fabs(h);
            double ynew[NV], fnew[NV];
            // error_norm^2; the controller never takes the square root:
            // error_norm < 1 <=> e2 < 1, error_norm^(-1/5) = e2^(-1/10)
            const double e2 = L.attempt(th, y, f, h, P.rtol, P.atol, ynew, fnew) * inv_n;
            const bool live = !done && !fail;
            const bool acc = live && (e2 < 1.0);
            const bool rej = live && !acc;   // NaN error norms land here too, as in scipy (nan < 1 is False)
            const double raw = rk_raw_factor<2>(e2);
            double f_acc = (e2 == 0.0) ? RK_MAX_FACTOR : fmin(RK_MAX_FACTOR, raw);
            if (step_rejected) f_acc = fmin(1.0, f_acc);
            const double f_rej = fmax(RK_MIN_FACTOR, raw);
            if (live) h_abs = h_abs_used * (acc ? f_acc : f_rej);
            if (acc) {
                t = t_new;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    y[i] = ynew[i];
                    f[i] = fnew[i];
                }
                add_moments();
                ++n_acc;
                new_step = true;
                if (t == P.T) done = true;
            }
            if (rej) {
                step_rejected = true;
                new_step = false;
                ++n_rej;
            }
            ++guard;
            if (fail || guard >= P.max_attempts) done = true;
        }
    }

    // ---- the same controller arranged for the dependent chain that links one attempt to the next
    // (error norm -> step factor -> h -> first stage).  A lone warp on a sub-partition cannot hide it, so:
    //  * the accept bookkeeping is branch-free selects (no divergent block of register moves between the
    //    factor and the next attempt: ptxas schedules them into the shadow of the power);
    //  * the clamps of the controller (factor <= 10, <= 1 after a rejection, >= 0.2; rk.py:157-171) are
    //    predicates on the squared error norm evaluated beside the power -- 0.9 x^(-1/10) >= c  <=>
    //    x <= (0.9/c)^10 -- and select the exact clamp values afterwards;
    //  * the power is rk_hfactor_fused (no F2F conversions, one 5-level Newton round, h folded into its last FMA);
    //  * the 1/n of the RMS norm is folded into the thresholds (sum of squares ss = n e2);
    //  * behind the power there is ONE select (every special case -- the clamps, a step below min_step -- is
    //    decided and its value formed beside the power), then the clip to t_bound as a compare of the new step
    //    with T - t (known since the accept select) and a select: the first stage of the next attempt follows the
    //    residual of the Newton round after FMA, select, DSETP, select.  scipy's h = (t + h_abs) - t
    //    (rk.py:137-141) is taken as h_abs itself (they differ by the rounding of t + h_abs, ~1e-13 relative);
    //  * a step below min_step (rk.py:122-131, never seen outside a failing integration) is not repaired in front
    //    of the attempt but costs one idle attempt: the lane sits the attempt out, takes min_step (new step) or
    //    fails (inside a step) afterwards -- the same sequence of steps as scipy, one pass later.
    template <bool C1>
    __device__ __forceinline__ void attempts_fused(const LorenzLanes<J, KT, NUM> &L, const LorenzDev &P, const LorenzTheta &th,
                                                   double atol_z) {
        const double n = (double)P.nvar;
        const double inv_n = 1.0 / n;
        const double ss_one = n;                              // error_norm < 1
        const double ss_cap10 = n * 3.486784401e-11;          // raw factor >= 10  <=> e2 <= 0.09^10
        const double ss_cap1 = n * 0.3486784401;              // raw factor >= 1   <=> e2 <= 0.9^10
        const double ss_floor = n * 3405062.8916015625;       // raw factor <= 0.2 <=> e2 >= 4.5^10
        const float log2n_10 = 0.1f * log2f((float)P.nvar);
        // Scheduling aid (no arithmetic effect): ptxas computes the accept predicate -- and with it the twenty accept
        // selects -- only where it is first needed, i.e. BEHIND the power, although it is known ~100 cycles earlier.
        // The seed of the power takes its bias from one of two identical registers picked by that predicate, which
        // puts the predicate at the head of the critical chain and the selects into the idle slots of the power.
        // (the twin differs only for a negative attempt budget, which never runs: a run-time fact ptxas cannot fold)
        const float log2n_10_twin = log2n_10 + (P.max_attempts < 0 ? 1.0f : 0.0f);
        int guard = 0;
        // the distance to t_bound depends on t alone: carried across the back edge and recomputed as soon as the
        // accept select has produced t
        double rem = P.T - t;
        // the attempt in flight: its step and the state of its second stage (rotated loop)
        double h = (h_abs > rem) ? rem : h_abs;               // t + h_abs beyond t_bound (rk.py:137-140); >= 0
        double ys2[NV];
        L.stage2(y, f, h, ys2);
        while (__any_sync(FULL, !done)) {
            double ynew[NV], fnew[NV];
            const double ss = L.template attempt_fused<C1>(th, y, f, ys2, h, P.rtol, P.atol, atol_z, ynew, fnew);
            // flags and end time of this attempt (in source order behind it so that they share its basic block --
            // the first shuffle of the loop body is preceded by a convergence check that ends the block at the loop top)
            // min_step = 10 |nextafter(t) - t| (rk.py:122; t >= 0: no fabs)
            const double min_step = 10.0 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
            const bool tiny = h_abs < min_step;               // off the chain: decides `live`
            const double t_sum = t + h_abs;
            const bool last = (h_abs > rem) || !(t_sum < P.T);   // this attempt ends at t_bound
            const double t_new = last ? P.T : t_sum;
            const bool live = !done && !tiny;
            const bool acc = live && (ss < ss_one);
            const bool rej = live && !acc;
            // next step size: raw power beside its clamp predicates (NaN: no predicate holds except the floor)
            const bool at_cap = ss <= (step_rejected ? ss_cap1 : ss_cap10);   // incl. ss == 0 (rk.py:157-160)
            const bool at_floor = !(ss < ss_floor);                            // incl. NaN (max(0.2, nan) = 0.2)
            const double clamp_fac = at_floor ? RK_MIN_FACTOR : (step_rejected ? 1.0 : RK_MAX_FACTOR);
            const double h_alt = tiny ? min_step : h * clamp_fac;
            const bool use_alt = tiny || at_cap || at_floor;
            const double h_raw = rk_hfactor_fused(ss, inv_n, acc ? log2n_10 : log2n_10_twin, h);
            h_abs = use_alt ? h_alt : h_raw;
            // accept bookkeeping as selects
            t = acc ? t_new : t;
            rem = P.T - t;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                y[i] = acc ? ynew[i] : y[i];
                f[i] = acc ? fnew[i] : f[i];
            }
            if (acc) {                                                // predicated fp64 instructions, no selects
                const double X = ynew[0], Y0 = (J > 0) ? ynew[1] : 0.0;
                msum[0] += X;
                msum[1] += Y0;
                msum[2] = fma(X, X, msum[2]);
                msum[3] = fma(X, Y0, msum[3]);
                msum[4] = fma(Y0, Y0, msum[4]);
            }
            n_acc += acc ? 1 : 0;                             // n_t = n_acc + 1 (t0 is stored too) after the loop
            n_rej += rej ? 1 : 0;
            const bool fail = tiny && !new_step;               // TOO_SMALL_STEP inside a step: solve_ivp stops
            new_step = acc || (new_step && !rej);
            step_rejected = !new_step && (rej || step_rejected);   // a new step starts with the flag cleared
            ++guard;
            done = done || fail || guard >= P.max_attempts || (acc && last);
            // the next attempt: its step and second stage
            h = (h_abs > rem) ? rem : h_abs;
            L.stage2(y, f, h, ys2);
        }
    }
};

}  // namespace ipmcmc
