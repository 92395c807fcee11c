// Dynamic step scheduler shared by the Burgers and the Lorenz chain kernels: a FIFO of READY work
// units (a chain, or a warp's group of Lorenz chains) in global memory, served by persistent warps.
#pragma once
#include "common.cuh"

namespace ipmcmc {

// ------------------------------------------------------------------------------------------------
// Dynamic step scheduler.  Solve lengths are data dependent and the warps of an SM do not all run
// at the same speed (a warp alone on a sub-partition advances ~1.85x faster than one that shares
// it), so a static chain -> warp map leaves sub-partitions idle at the end of a launch.  Here
// persistent warps take (chain, `chunk` steps) work items from a FIFO of READY chains in global
// memory: a chain is pushed back as soon as its item is done, so chains rotate over the warps,
// advance at the same average pace and every warp stays busy until the queue runs dry.  The chain
// state (u, Phi, moments: ~100 B) travels through L2 between items; Philox is keyed by (chain,
// step), so the results do not depend on which warp ran which item (bit-identical to the static
// kernel; tested).
//
// Scratch (caller-owned, int64): ring[cap] | progress[n] | head | tail, cap = 2*n (n = work units).
// Ring entry = ((lap+1) << 32) | chain for the ticket lap*cap + slot; 0 = consumed/empty.
// ------------------------------------------------------------------------------------------------
struct SchedView {
    unsigned long long *ring, *head, *tail;
    long long *progress;
    long long cap;
    __device__ __forceinline__ SchedView(long long *base, long long n_chains)
        : ring((unsigned long long *)base), head((unsigned long long *)base + 3 * n_chains),
          tail((unsigned long long *)base + 3 * n_chains + 1), progress(base + 2 * n_chains), cap(2 * n_chains) {}
};
__host__ __device__ inline long long sched_len(long long n_chains) { return 3 * n_chains + 2; }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

static __global__ void sched_init_kernel(long long *sched, long long n_chains) {
    SchedView Q(sched, n_chains);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Q.cap; i += (long long)gridDim.x * blockDim.x) {
        Q.ring[i] = i < n_chains ? ((1ull << 32) | (unsigned long long)i) : 0ull;
        if (i < n_chains) Q.progress[i] = 0;
        if (i == 0) {
            *Q.head = 0;
            *Q.tail = (unsigned long long)n_chains;
        }
    }
}

// Spin until lane 0 reads a value accepted by `ok` from *p; returns it on every lane.  The loop
// condition is warp-uniform (lane 0 only issues a predicated load): a lane spinning on its own
// does not reconverge with the other 31, and the whole solve would then run twice, once per part
// of the split warp (measured: 4x slower).
template <class OK>
__device__ __forceinline__ unsigned long long spin_until(const unsigned long long *p, int lane, OK ok) {
    while (true) {
        unsigned long long v = 0;
        if (lane == 0) v = ld_acquire_u64(p);
        v = __shfl_sync(FULL, v, 0);
        if (ok(v)) return v;
        __nanosleep(64);
    }
}


}  // namespace ipmcmc
