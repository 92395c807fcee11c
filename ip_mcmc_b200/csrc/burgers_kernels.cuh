// Burgers kernels: batched forward evaluation and the fused Metropolis kernel.
// One chain per warp, one warp per CTA (the hardware block scheduler then balances the
// data-dependent solve lengths across the 148 SMs); the FV state never leaves registers, the end
// state is staged once per solve in shared memory for the windowed trapezoid measurement.
#pragma once
#include "burgers.cuh"
#include "burgers_team.cuh"
#include "sampler.cuh"
#include "sched.cuh"

// Register budget of the static chain kernel: 128 registers (2 CTAs of 256 threads per SM) up to 8 cells
// per lane, 255 from 16 cells per lane on, where 128 spill the state (measured: engine.cu,
// burgers_launch_chain_queue).
#ifndef IPMCMC_CHAIN_MINB
#define IPMCMC_CHAIN_MINB(CPL) ((CPL) >= 16 ? 1 : 2)
#endif

namespace ipmcmc {

// Developer instrumentation (-DIPMCMC_PROF=1, tools/overhead_probe.py): cycles per section of a work
// item, summed over all warps by lane 0.  Not compiled into the product library.
#ifndef IPMCMC_PROF
#define IPMCMC_PROF 0
#endif
#if IPMCMC_PROF
static __device__ unsigned long long g_prof[16];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, a, b) do { if (lane_id() == 0) atomicAdd(&g_prof[slot], (unsigned long long)((b) - (a))); } while (0)
#else
#define PROF_T(var)
#define PROF_ADD(slot, a, b)
#endif

// shared memory layout per warp: state[N] | G[MAX_OBS] | r2[MAX_OBS]
__host__ __device__ inline size_t burgers_smem_bytes(int N, int warps = 1) {
    return (size_t)warps * (N + 2 * IPMCMC_MAX_OBS) * sizeof(double);
}

// Measurer.__call__ (utilities.py:100-109): m_i = 10 * trapz(values[l:r], dx) with numpy's
// evaluation (dx*(y[1:]+y[:-1])/2.0).sum() in pairwise order; lane i handles window i.
__device__ __forceinline__ void burgers_measure(const BurgersDev &B, const double *state, double *Gs, int lane) {
    for (int i = lane; i < B.pot.q; i += 32) {
        const int l = B.win_left[i], r = B.win_right[i];
        const int nterm = r - l - 1;
        double sum = 0.0;
        if (nterm >= 1) {
            const double dxm = B.dx_meas;
            sum = np_pairwise_sum([&](int j) { return dxm * (state[j + 1] + state[j]) / 2.0; }, l, nterm);
        }
        Gs[i] = 10.0 * sum;
    }
    __syncwarp();
}

// G(u) and Phi(u) for the parameter vector whose component i sits on lane i (value `ui`).
// Leaves the end state in smem `state` and G in `Gs`.  Returns Phi; n_fv by reference.
// Inlined at its call sites on purpose: with the solve as an out-of-line subroutine ptxas allocates the
// time-step loop's registers around the call convention and schedules it measurably worse (B200,
// 1024 x 256 cells: 60 % of the fp64 peak out of line, 66 % inlined; 8192 x 1024: 78 % / 83 %).
#ifndef IPMCMC_PHI_INLINE
#define IPMCMC_PHI_INLINE __forceinline__
#endif
template <int CPL, int NUMERICS, bool PADDED>
__device__ IPMCMC_PHI_INLINE double burgers_phi(const BurgersDev &B, double ui, double *state, double *Gs, double *r2,
                                              int lane, int &n_fv) {
    // FVMObservationOperator.__call__ (utilities.py:40-41): IC(u_0 + u)
    const double pi = (lane < B.d) ? B.param_mean[lane] + ui : 0.0;
    BurgersWarp<CPL, NUMERICS, PADDED> W;
    PROF_T(p0);
    n_fv = W.integrate(B, pi, lane);
    PROF_T(p1);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int c = lane * CPL + k;
        if (c < B.N) state[c] = W.u[k];
    }
    __syncwarp();
    burgers_measure(B, state, Gs, lane);
    const double phi = potential_from_G(B.pot, Gs, r2, lane, 32, FULL);
    PROF_T(p2);
    PROF_ADD(3, p0, p1);
    PROF_ADD(4, p1, p2);
    // a solve stopped by the safety cap has not reached T: report it as non-finite (-> rejected)
    return W.capped ? nan("") : phi;
}

template <int CPL, int NUMERICS, bool PADDED>
__global__ void __launch_bounds__(256) burgers_forward_kernel(const __grid_constant__ BurgersDev B, long long n,
                                                              const double *__restrict__ u, double *__restrict__ G,
                                                              double *__restrict__ phi, double *__restrict__ state_out,
                                                              long long *__restrict__ work) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    double *smem = smem_all + (size_t)warp * (B.N + 2 * IPMCMC_MAX_OBS);
    double *state = smem, *Gs = smem + B.N, *r2 = Gs + IPMCMC_MAX_OBS;
    const int lane = lane_id();
    for (long long c = (long long)blockIdx.x * wpc + warp; c < n; c += (long long)gridDim.x * wpc) {
        const double ui = (lane < B.d) ? u[c * B.d + lane] : 0.0;
        int n_fv;
        const double ph = burgers_phi<CPL, NUMERICS, PADDED>(B, ui, state, Gs, r2, lane, n_fv);
        if (G)
            for (int i = lane; i < B.pot.q; i += 32) G[c * B.pot.q + i] = Gs[i];
        if (state_out)
            for (int i = lane; i < B.N; i += 32) state_out[c * B.N + i] = state[i];
        if (lane == 0) {
            if (phi) phi[c] = ph;
            if (work) {
                work[2 * c] = n_fv;
                work[2 * c + 1] = 0;
            }
        }
        __syncwarp();
    }
}

// Registers of one chain between Metropolis steps.
struct ChainRegs {
    double ui, phi_u, reg_u;
    Welford mom;
    long long cnt[CNT_N];
};

// ONE Metropolis step of chain c (local id; cg global id) at launch-local step s: proposal -> (box
// constraint) -> solve -> Phi(v) -> accept/reject -> counters, logs, moments, trace.
// `phi_of(ui, n_fv)` evaluates Phi for the parameter vector whose component i sits on lane i: the one-warp
// solver (burgers_phi) or the multi-warp team solver (burgers_team_phi), whose warps all run this same step
// with the same Philox draws; `writer` (warp 0 of a team; always true for one warp per chain) owns the
// global-memory outputs.
#ifndef IPMCMC_STEP_INLINE
#define IPMCMC_STEP_INLINE __forceinline__
#endif
template <class PhiOf>
__device__ IPMCMC_STEP_INLINE void metropolis_step(const SamplerDev &S, const ChainBufDev &C, const Group &Gp, long long c,
                                                   long long cg, long long s, long long n_steps, ChainRegs &R,
                                                   const PhiOf &phi_of, bool writer) {
    const int lane = Gp.lane, d = S.d;
    const long long gstep = S.first_step + s;
    double ca, cb;
    PROF_T(q0);
    step_coefs(S, gstep, ca, cb);
    const double w = proposal_noise(S, C, Gp, c, cg, s, n_steps, gstep);
    const double vi = ca * R.ui + cb * w;
    PROF_T(q1);
    PROF_ADD(2, q0, q1);
    if (writer && C.vlog && lane < d) C.vlog[(c * n_steps + s) * d + lane] = vi;
    bool accepted = false;
    double phi_v = nan(""), a = nan("");
    int n_fv = 0;
    const bool ok = !S.has_constraint || constraint_ok(S, Gp, vi);
    if (ok) {
        if (S.recompute_phi_u) {  // the reference's 2 solves per step (accepter.py:121-122)
            int nf0;
            R.phi_u = phi_of(R.ui, nf0);
            R.cnt[CNT_WORK_A] += nf0;
            R.cnt[CNT_WORK_B] += 1;
        }
        phi_v = phi_of(vi, n_fv);
        PROF_T(q2);
        R.cnt[CNT_WORK_A] += n_fv;
        R.cnt[CNT_WORK_B] += 1;
        double reg_v = 0.0;
        if (S.accepter == IPMCMC_ACCEPT_RW) reg_v = prior_regulariser(S, Gp, vi);
        a = exp((R.phi_u + R.reg_u) - (phi_v + reg_v));
        const double U = C.inject_u ? C.inject_u[c * n_steps + s] : draw_uniform(S.seed, (uint64_t)cg, (uint64_t)gstep);
        accepted = a > U;  // strict, un-clipped; NaN compares false (accepter.py:61-62)
        if (!isfinite(phi_v)) R.cnt[CNT_NONFINITE] += 1;
        if (accepted) {
            R.ui = vi;
            R.phi_u = phi_v;
            R.reg_u = reg_v;
        }
        PROF_T(q3);
        PROF_ADD(5, q2, q3);
    } else {
        R.cnt[CNT_CONSTRAINT] += 1;
    }
    R.cnt[CNT_CALLS] += 1;
    R.cnt[CNT_ACCEPTS] += accepted ? 1 : 0;
    if (writer && C.steplog && lane == 0) {
        double *L = C.steplog + (c * n_steps + s) * 4;
        L[0] = phi_v;
        L[1] = a;
        L[2] = accepted ? 1.0 : 0.0;
        L[3] = (double)n_fv;
    }
    // recording (sampler.py:23-28)
    if (records_step(S, gstep)) {
        R.mom.add(R.ui);
        const long long n_rec = recorded_before(S, gstep) - recorded_before(S, S.first_step);
        if (writer && C.trace && n_rec < C.n_record && lane < d) C.trace[(c * C.n_record + n_rec) * d + lane] = R.ui;
    }
}

template <int CPL, int NUMERICS, bool PADDED>
__device__ IPMCMC_STEP_INLINE void burgers_metropolis_step(const BurgersDev &B, const SamplerDev &S, const ChainBufDev &C,
                                                        const Group &Gp, long long c, long long cg, long long s,
                                                        long long n_steps, double *state, double *Gs, double *r2,
                                                        ChainRegs &R) {
    const int lane = Gp.lane;
    metropolis_step(S, C, Gp, c, cg, s, n_steps, R,
                    [&](double ui, int &n_fv) { return burgers_phi<CPL, NUMERICS, PADDED>(B, ui, state, Gs, r2, lane, n_fv); },
                    true);
}

// W warps per CTA, one chain per warp.  Warps never synchronise with each other; the CTA shape
// only pins which chains share an SM sub-partition (warp w -> SMSP w % 4).
template <int CPL, int NUMERICS, bool PADDED>
__global__ void __launch_bounds__(256, IPMCMC_CHAIN_MINB(CPL)) burgers_chain_kernel(const __grid_constant__ BurgersDev B,
                                                            const __grid_constant__ SamplerDev S,
                                                            const __grid_constant__ ChainBufDev C, long long n_chains,
                                                            long long n_steps) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    double *smem = smem_all + (size_t)warp * (B.N + 2 * IPMCMC_MAX_OBS);
    double *state = smem, *Gs = smem + B.N, *r2 = Gs + IPMCMC_MAX_OBS;
    const int lane = lane_id();
    const Group Gp{0, 32, lane, FULL};
    const int d = S.d;
    const long long n_slots = C.slot_chain ? (long long)C.n_slots : n_chains;
    for (long long slot = (long long)blockIdx.x * wpc + warp; slot < n_slots; slot += (long long)gridDim.x * wpc) {
        const long long c = C.slot_chain ? (long long)C.slot_chain[slot] : slot;
        if (c < 0 || c >= n_chains) continue;
        const long long cg = S.chain_offset + c;
        ChainRegs R;
        R.ui = (lane < d) ? C.u[c * d + lane] : 0.0;
        R.phi_u = C.phi[c];
#pragma unroll
        for (int k = 0; k < CNT_N; ++k) R.cnt[k] = 0;
        if (isnan(R.phi_u)) {  // first launch: Phi(u_0) not known yet
            int n_fv;
            R.phi_u = burgers_phi<CPL, NUMERICS, PADDED>(B, R.ui, state, Gs, r2, lane, n_fv);
            R.cnt[CNT_WORK_A] += n_fv;
            R.cnt[CNT_WORK_B] += 1;
        }
        R.reg_u = (S.accepter == IPMCMC_ACCEPT_RW) ? prior_regulariser(S, Gp, R.ui) : 0.0;
        R.mom = Welford{C.mom_count[c], (lane < d) ? C.mom_mean[c * d + lane] : 0.0,
                        (lane < d) ? C.mom_m2[c * d + lane] : 0.0};
        for (long long s = 0; s < n_steps; ++s)
            burgers_metropolis_step<CPL, NUMERICS, PADDED>(B, S, C, Gp, c, cg, s, n_steps, state, Gs, r2, R);
        // write back
        if (lane < d) {
            C.u[c * d + lane] = R.ui;
            C.mom_mean[c * d + lane] = R.mom.mean;
            C.mom_m2[c * d + lane] = R.mom.m2;
        }
        if (lane == 0) {
            C.phi[c] = R.phi_u;
            C.mom_count[c] = R.mom.count;
#pragma unroll
            for (int k = 0; k < CNT_N; ++k) C.counters[c * CNT_N + k] += R.cnt[k];
        }
        __syncwarp();
    }
}

// MINB: register budget, 2 = 128 registers (16 warps per SM), 1 = 255 registers; chosen by the number of
// cells per lane in burgers_launch_chain_queue (engine.cu), where the measurements are quoted.
template <int CPL, int NUMERICS, bool PADDED, int MINB>
__global__ void __launch_bounds__(256, MINB) burgers_chain_queue_kernel(
    const __grid_constant__ BurgersDev B, const __grid_constant__ SamplerDev S, const __grid_constant__ ChainBufDev C,
    long long n_chains, long long n_steps, int chunk) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5;
    double *smem = smem_all + (size_t)warp * (B.N + 2 * IPMCMC_MAX_OBS);
    double *state = smem, *Gs = smem + B.N, *r2 = Gs + IPMCMC_MAX_OBS;
    const int lane = lane_id();
    const Group Gp{0, 32, lane, FULL};
    const int d = S.d;
    SchedView Q(C.sched, n_chains);
    const unsigned long long cap = (unsigned long long)Q.cap;
    const long long items_per_chain = (n_steps + chunk - 1) / chunk;
    const unsigned long long total = (unsigned long long)n_chains * (unsigned long long)items_per_chain;
    while (true) {
        // ---- take a ticket, wait for its ring slot to be filled, consume it
        PROF_T(t0);
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(Q.head, 1ull);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= total) break;
        unsigned long long *slot = Q.ring + idx % cap;
        const unsigned long long want = idx / cap + 1;
        const unsigned long long e = spin_until(slot, lane, [want](unsigned long long v) { return (v >> 32) == want; });
        // mark the slot consumed: ordered after our own read of it (same address), nothing to publish
        if (lane == 0) st_relaxed_u64(slot, 0ull);
        __syncwarp();
        const long long c = (long long)(e & 0xffffffffull), cg = S.chain_offset + c;
        // ---- chain state from L2 (another SM may have written it: bypass L1)
        const long long s0 = __ldcg(Q.progress + c);
        ChainRegs R;
        R.ui = (lane < d) ? __ldcg(C.u + c * d + lane) : 0.0;
        R.phi_u = __ldcg(C.phi + c);
#pragma unroll
        for (int k = 0; k < CNT_N; ++k) R.cnt[k] = 0;
        if (isnan(R.phi_u)) {  // first launch: Phi(u_0) not known yet
            int n_fv;
            R.phi_u = burgers_phi<CPL, NUMERICS, PADDED>(B, R.ui, state, Gs, r2, lane, n_fv);
            R.cnt[CNT_WORK_A] += n_fv;
            R.cnt[CNT_WORK_B] += 1;
        }
        R.reg_u = (S.accepter == IPMCMC_ACCEPT_RW) ? prior_regulariser(S, Gp, R.ui) : 0.0;
        R.mom = Welford{__ldcg(C.mom_count + c), (lane < d) ? __ldcg(C.mom_mean + c * d + lane) : 0.0,
                        (lane < d) ? __ldcg(C.mom_m2 + c * d + lane) : 0.0};
        const long long s1 = (s0 + chunk < n_steps) ? s0 + chunk : n_steps;
        PROF_T(t1);
        for (long long s = s0; s < s1; ++s)
            burgers_metropolis_step<CPL, NUMERICS, PADDED>(B, S, C, Gp, c, cg, s, n_steps, state, Gs, r2, R);
        PROF_T(t2);
        // ---- write back, then hand the chain to the next free warp.  The chain state is published by
        // the release store of the ring entry (no separate fence): the other lanes' stores are ordered
        // before it by __syncwarp (cumulativity).
        if (lane < d) {
            __stcg(C.u + c * d + lane, R.ui);
            __stcg(C.mom_mean + c * d + lane, R.mom.mean);
            __stcg(C.mom_m2 + c * d + lane, R.mom.m2);
        }
        if (lane == 0) {
            __stcg(C.phi + c, R.phi_u);
            __stcg(C.mom_count + c, R.mom.count);
#pragma unroll
            for (int k = 0; k < CNT_N; ++k) atomicAdd((unsigned long long *)(C.counters + c * CNT_N + k), (unsigned long long)R.cnt[k]);
            __stcg(Q.progress + c, s1);
        }
        __syncwarp();
        if (s1 < n_steps) {   // warp-uniform
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(Q.tail, 1ull);
            t = __shfl_sync(FULL, t, 0);
            unsigned long long *pslot = Q.ring + t % cap;
            spin_until(pslot, lane, [](unsigned long long v) { return v == 0ull; });   // previous lap consumed (cap = 2n: no wait in practice)
            if (lane == 0) st_release_u64(pslot, ((t / cap + 1) << 32) | (unsigned long long)c);
        }
        __syncwarp();
        PROF_T(t3);
        PROF_ADD(0, t0, t3);
        PROF_ADD(1, t0, t1);
        PROF_ADD(6, t1, t2);
        PROF_ADD(7, t2, t3);
        PROF_ADD(8, 0, 1);
    }
}

// ------------------------------------------------------------------------------------------------
// wide parameter vectors (IPMCMC_MAX_DIM < d <= IPMCMC_MAX_DIM_WIDE): the truncated KL / spectral prior
// with up to 253 modes.  One chain per warp as before, but the parameter vector no longer fits the lanes
// of the warp: u, the proposal v and the absolute parameters live in shared memory (component i is served
// by lane i % 32), the solver reads its parameters from there (BurgersWarp::integrate over a ParamAt
// functor), the running moments are updated in place in global memory.  Same Philox draws (slot = component),
// same accept rule; diagonal sampling factor and diagonal prior factor only (what a KL prior is); static
// chain -> warp map.  CPU statement: oracle/mcmc_np.run_chain + oracle/burgers_np with kl_basis.
// ------------------------------------------------------------------------------------------------
// shared memory per warp: state[N] | G[MAX_OBS] | r2[MAX_OBS] | par[d] | u[d] | v[d]
__host__ __device__ inline size_t burgers_wide_smem_bytes(int N, int d, int warps = 1) {
    return (size_t)warps * (N + 2 * IPMCMC_MAX_OBS + 3 * d) * sizeof(double);
}

// Phi for the parameter perturbation x[0..d) (shared memory); par[] is scratch for mean + x
template <int CPL, int NUMERICS, bool PADDED>
__device__ __forceinline__ double burgers_phi_wide(const BurgersDev &B, const double *x, double *par, double *state,
                                                   double *Gs, double *r2, int lane, int &n_fv) {
    for (int i = lane; i < B.d; i += 32) par[i] = B.param_mean_wide[i] + x[i];   // utilities.py:41
    __syncwarp();
    BurgersWarp<CPL, NUMERICS, PADDED> W;
    n_fv = W.integrate(B, [par](int i) { return par[i]; }, lane);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int c = lane * CPL + k;
        if (c < B.N) state[c] = W.u[k];
    }
    __syncwarp();
    burgers_measure(B, state, Gs, lane);
    const double phi = potential_from_G(B.pot, Gs, r2, lane, 32, FULL);
    return W.capped ? nan("") : phi;
}

// sum over the warp in a fixed order (lane tree), the same bits on every lane
__device__ __forceinline__ double warp_sum_fixed(double v) {
#pragma unroll
    for (int off = 1; off < 32; off *= 2) {
        const double o = __shfl_xor_sync(FULL, v, off);
        v = (lane_id() & off) ? o + v : v + o;
    }
    return v;
}

// 0.5 * ||L x||^2 for a diagonal L (accepter.py:104-106)
__device__ __forceinline__ double prior_regulariser_wide(const SamplerDev &S, const double *x, int lane) {
    double ss = 0.0;
    for (int i = lane; i < S.d; i += 32) {
        const double y = S.prior_chol_diag[i] * x[i];
        ss = ss + y * y;
    }
    const double nrm = sqrt(warp_sum_fixed(ss));
    return 0.5 * (nrm * nrm);
}

template <int CPL, int NUMERICS, bool PADDED>
__global__ void __launch_bounds__(128) burgers_wide_forward_kernel(const __grid_constant__ BurgersDev B, long long n,
                                                                   const double *__restrict__ u, double *__restrict__ G,
                                                                   double *__restrict__ phi, double *__restrict__ state_out,
                                                                   long long *__restrict__ work) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5, d = B.d;
    double *smem = smem_all + (size_t)warp * (B.N + 2 * IPMCMC_MAX_OBS + 3 * d);
    double *state = smem, *Gs = smem + B.N, *r2 = Gs + IPMCMC_MAX_OBS, *par = r2 + IPMCMC_MAX_OBS, *x = par + d;
    const int lane = lane_id();
    for (long long c = (long long)blockIdx.x * wpc + warp; c < n; c += (long long)gridDim.x * wpc) {
        for (int i = lane; i < d; i += 32) x[i] = u[c * d + i];
        __syncwarp();
        int n_fv;
        const double ph = burgers_phi_wide<CPL, NUMERICS, PADDED>(B, x, par, state, Gs, r2, lane, n_fv);
        if (G)
            for (int i = lane; i < B.pot.q; i += 32) G[c * B.pot.q + i] = Gs[i];
        if (state_out)
            for (int i = lane; i < B.N; i += 32) state_out[c * B.N + i] = state[i];
        if (lane == 0) {
            if (phi) phi[c] = ph;
            if (work) {
                work[2 * c] = n_fv;
                work[2 * c + 1] = 0;
            }
        }
        __syncwarp();
    }
}

template <int CPL, int NUMERICS, bool PADDED>
__global__ void __launch_bounds__(128) burgers_wide_chain_kernel(const __grid_constant__ BurgersDev B,
                                                                 const __grid_constant__ SamplerDev S,
                                                                 const __grid_constant__ ChainBufDev C, long long n_chains,
                                                                 long long n_steps) {
    extern __shared__ double smem_all[];
    const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5, d = S.d;
    double *smem = smem_all + (size_t)warp * (B.N + 2 * IPMCMC_MAX_OBS + 3 * d);
    double *state = smem, *Gs = smem + B.N, *r2 = Gs + IPMCMC_MAX_OBS, *par = r2 + IPMCMC_MAX_OBS, *us = par + d, *vs = us + d;
    const int lane = lane_id();
    for (long long c = (long long)blockIdx.x * wpc + warp; c < n_chains; c += (long long)gridDim.x * wpc) {
        const long long cg = S.chain_offset + c;
        for (int i = lane; i < d; i += 32) us[i] = C.u[c * d + i];
        __syncwarp();
        double phi_u = C.phi[c];
        long long cnt[CNT_N];
#pragma unroll
        for (int k = 0; k < CNT_N; ++k) cnt[k] = 0;
        if (isnan(phi_u)) {  // first launch: Phi(u_0) not known yet
            int n_fv;
            phi_u = burgers_phi_wide<CPL, NUMERICS, PADDED>(B, us, par, state, Gs, r2, lane, n_fv);
            cnt[CNT_WORK_A] += n_fv;
            cnt[CNT_WORK_B] += 1;
        }
        double reg_u = (S.accepter == IPMCMC_ACCEPT_RW) ? prior_regulariser_wide(S, us, lane) : 0.0;
        double count = C.mom_count[c];
        for (long long s = 0; s < n_steps; ++s) {
            const long long gstep = S.first_step + s;
            double ca, cb;
            step_coefs(S, gstep, ca, cb);
            // proposal (proposer.py:29-30, 81-82): v = ca*u + cb*w, w_i = factor_i * z_i, z_i = Philox slot i
            bool inside = true;
            for (int i = lane; i < d; i += 32) {
                double w;
                if (C.inject_w) {
                    w = C.inject_w[(c * n_steps + s) * d + i];
                } else {
                    const double z = draw_normal(S.seed, (uint64_t)cg, (uint64_t)gstep, (uint32_t)i);
                    w = S.factor_kind == 1 ? S.factor[i] * z : z;
                }
                const double vi = ca * us[i] + cb * w;
                vs[i] = vi;
                if (C.vlog) C.vlog[(c * n_steps + s) * d + i] = vi;
                if (S.has_constraint) {
                    const double sh = vi + S.box_wide[2 * d + i];
                    inside = inside && (sh > S.box_wide[i]) && (sh < S.box_wide[d + i]);
                }
            }
            __syncwarp();
            const bool ok = !S.has_constraint || __all_sync(FULL, inside);
            bool accepted = false;
            double phi_v = nan(""), a = nan("");
            int n_fv = 0;
            if (ok) {
                if (S.recompute_phi_u) {  // the reference's 2 solves per step (accepter.py:121-122)
                    int nf0;
                    phi_u = burgers_phi_wide<CPL, NUMERICS, PADDED>(B, us, par, state, Gs, r2, lane, nf0);
                    cnt[CNT_WORK_A] += nf0;
                    cnt[CNT_WORK_B] += 1;
                }
                phi_v = burgers_phi_wide<CPL, NUMERICS, PADDED>(B, vs, par, state, Gs, r2, lane, n_fv);
                cnt[CNT_WORK_A] += n_fv;
                cnt[CNT_WORK_B] += 1;
                const double reg_v = (S.accepter == IPMCMC_ACCEPT_RW) ? prior_regulariser_wide(S, vs, lane) : 0.0;
                a = exp((phi_u + reg_u) - (phi_v + reg_v));
                const double U = C.inject_u ? C.inject_u[c * n_steps + s] : draw_uniform(S.seed, (uint64_t)cg, (uint64_t)gstep);
                accepted = a > U;  // strict, un-clipped; NaN compares false (accepter.py:61-62)
                if (!isfinite(phi_v)) cnt[CNT_NONFINITE] += 1;
                if (accepted) {
                    for (int i = lane; i < d; i += 32) us[i] = vs[i];
                    phi_u = phi_v;
                    reg_u = reg_v;
                }
                __syncwarp();
            } else {
                cnt[CNT_CONSTRAINT] += 1;
            }
            cnt[CNT_CALLS] += 1;
            cnt[CNT_ACCEPTS] += accepted ? 1 : 0;
            if (C.steplog && lane == 0) {
                double *L = C.steplog + (c * n_steps + s) * 4;
                L[0] = phi_v;
                L[1] = a;
                L[2] = accepted ? 1.0 : 0.0;
                L[3] = (double)n_fv;
            }
            if (records_step(S, gstep)) {   // Welford in place (sampler.py:23-28)
                count += 1.0;
                const long long n_rec = recorded_before(S, gstep) - recorded_before(S, S.first_step);
                for (int i = lane; i < d; i += 32) {
                    const double x = us[i], m0 = C.mom_mean[c * d + i];
                    const double delta = x - m0, m1 = m0 + delta / count;
                    C.mom_mean[c * d + i] = m1;
                    C.mom_m2[c * d + i] += delta * (x - m1);
                    if (C.trace && n_rec < C.n_record) C.trace[(c * C.n_record + n_rec) * d + i] = x;
                }
            }
        }
        for (int i = lane; i < d; i += 32) C.u[c * d + i] = us[i];
        if (lane == 0) {
            C.phi[c] = phi_u;
            C.mom_count[c] = count;
#pragma unroll
            for (int k = 0; k < CNT_N; ++k) C.counters[c * CNT_N + k] += cnt[k];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// team kernels: one CTA of TM warps per chain (N = TM*32*CPL cells > 1024)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t burgers_team_smem_bytes(int N) {
    return sizeof(TeamXch) + (size_t)(N + 2 * IPMCMC_MAX_OBS) * sizeof(double);
}

// Every warp of the team calls this with the same ui; warp 0 measures and evaluates Phi, the value is
// broadcast through shared memory so that all warps take the same accept decision.
template <int CPL, int NUMERICS, int TM>
__device__ __noinline__ double burgers_team_phi(const BurgersDev &B, TeamXch &X, double ui, double *state, double *Gs,
                                                double *r2, int tw, int lane, int &n_fv) {
    const double pi = (lane < B.d) ? B.param_mean[lane] + ui : 0.0;
    BurgersTeam<CPL, NUMERICS, TM> W;
    n_fv = W.integrate(B, X, pi, tw, lane);
#pragma unroll
    for (int k = 0; k < CPL; ++k) state[(tw * 32 + lane) * CPL + k] = W.u[k];
    __syncthreads();
    if (tw == 0) {
        burgers_measure(B, state, Gs, lane);
        const double phi = potential_from_G(B.pot, Gs, r2, lane, 32, FULL);
        if (lane == 0) X.phi = W.capped ? nan("") : phi;
    }
    __syncthreads();
    return X.phi;
}

template <int CPL, int NUMERICS, int TM>
__global__ void __launch_bounds__(32 * TM) burgers_team_forward_kernel(const __grid_constant__ BurgersDev B, long long n,
                                                                      const double *__restrict__ u,
                                                                      double *__restrict__ G, double *__restrict__ phi,
                                                                      double *__restrict__ state_out,
                                                                      long long *__restrict__ work) {
    extern __shared__ double smem_all[];
    TeamXch &X = *reinterpret_cast<TeamXch *>(smem_all);
    double *state = smem_all + sizeof(TeamXch) / sizeof(double), *Gs = state + B.N, *r2 = Gs + IPMCMC_MAX_OBS;
    const int tw = threadIdx.x >> 5, lane = lane_id();
    for (long long c = blockIdx.x; c < n; c += gridDim.x) {
        const double ui = (lane < B.d) ? u[c * B.d + lane] : 0.0;
        int n_fv;
        const double ph = burgers_team_phi<CPL, NUMERICS, TM>(B, X, ui, state, Gs, r2, tw, lane, n_fv);
        if (G && tw == 0)
            for (int i = lane; i < B.pot.q; i += 32) G[c * B.pot.q + i] = Gs[i];
        if (state_out)
            for (int i = threadIdx.x; i < B.N; i += blockDim.x) state_out[c * B.N + i] = state[i];
        if (threadIdx.x == 0) {
            if (phi) phi[c] = ph;
            if (work) {
                work[2 * c] = n_fv;
                work[2 * c + 1] = 0;
            }
        }
        __syncthreads();
    }
}

template <int CPL, int NUMERICS, int TM>
__global__ void __launch_bounds__(32 * TM) burgers_team_chain_kernel(const __grid_constant__ BurgersDev B,
                                                                    const __grid_constant__ SamplerDev S,
                                                                    const __grid_constant__ ChainBufDev C,
                                                                    long long n_chains, long long n_steps) {
    extern __shared__ double smem_all[];
    TeamXch &X = *reinterpret_cast<TeamXch *>(smem_all);
    double *state = smem_all + sizeof(TeamXch) / sizeof(double), *Gs = state + B.N, *r2 = Gs + IPMCMC_MAX_OBS;
    const int tw = threadIdx.x >> 5, lane = lane_id();
    const bool writer = tw == 0;  // all warps run the same Metropolis logic; warp 0 owns the global writes
    const Group Gp{0, 32, lane, FULL};
    const int d = S.d;
    const auto phi_of = [&](double ui, int &n_fv) {
        return burgers_team_phi<CPL, NUMERICS, TM>(B, X, ui, state, Gs, r2, tw, lane, n_fv);
    };
    for (long long c = blockIdx.x; c < n_chains; c += gridDim.x) {
        const long long cg = S.chain_offset + c;
        ChainRegs R;
        R.ui = (lane < d) ? C.u[c * d + lane] : 0.0;
        R.phi_u = C.phi[c];
#pragma unroll
        for (int k = 0; k < CNT_N; ++k) R.cnt[k] = 0;
        __syncthreads();  // everyone has read the chain state before warp 0 may overwrite it at the end
        if (isnan(R.phi_u)) {
            int n_fv;
            R.phi_u = phi_of(R.ui, n_fv);
            R.cnt[CNT_WORK_A] += n_fv;
            R.cnt[CNT_WORK_B] += 1;
        }
        R.reg_u = (S.accepter == IPMCMC_ACCEPT_RW) ? prior_regulariser(S, Gp, R.ui) : 0.0;
        R.mom = Welford{C.mom_count[c], (lane < d) ? C.mom_mean[c * d + lane] : 0.0,
                        (lane < d) ? C.mom_m2[c * d + lane] : 0.0};
        for (long long s = 0; s < n_steps; ++s) metropolis_step(S, C, Gp, c, cg, s, n_steps, R, phi_of, writer);
        __syncthreads();
        if (writer) {
            if (lane < d) {
                C.u[c * d + lane] = R.ui;
                C.mom_mean[c * d + lane] = R.mom.mean;
                C.mom_m2[c * d + lane] = R.mom.m2;
            }
            if (lane == 0) {
                C.phi[c] = R.phi_u;
                C.mom_count[c] = R.mom.count;
#pragma unroll
                for (int k = 0; k < CNT_N; ++k) C.counters[c * CNT_N + k] += R.cnt[k];
            }
        }
        __syncthreads();
    }
}

}  // namespace ipmcmc
