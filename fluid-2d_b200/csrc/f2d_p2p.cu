// f2d_p2p.cu -- halo exchange by direct stores into the neighbour GPU's memory (NVLink peer access).
//
// The row-slab path first used NCCL send/recv pairs (f2d_solver.cu, still available).  On 8 B200s the
// ~12 exchanges per step cost ~0.11 ms each there (launch + rendezvous), 17 % of an 7.7 ms step, although
// only ~2 MB per neighbour move.  This transport does one kernel per exchange on the solver's stream:
//
//   1. announce READY(e) in both neighbours' flag blocks: "everything I launched before exchange e is
//      done, my halo rows may be overwritten" (stream order makes that true when the kernel starts);
//   2. wait for READY(e) from both neighbours (they are at the same point of the same schedule);
//   3. push my first/last owned rows straight into the neighbours' halo rows (16-byte peer stores);
//   4. fence (system scope), the last block announces DONE(e) to both neighbours, waits for their DONE(e)
//      (my own halos are then complete) and advances the epoch.
//
// All flags are monotonically increasing epochs in device memory, so the same kernels replay inside a
// CUDA graph step after step.  Every wait is bounded in wall-clock time (default 30 s, F2D_P2P_TIMEOUT_MS): on a
// timeout the error flag is raised (reported and cleared by f2d_sync) and the kernel exits instead of hanging the GPU.
#include "f2d_kernels.cuh"

namespace f2d {

namespace {
// flag block layout (unsigned words) inside every rank's arena
enum { FL_EPOCH = 0, FL_BLOCKS = 1, FL_ERROR = 2, FL_READY_FROM_UP = 8, FL_READY_FROM_DOWN = 9, FL_DONE_FROM_UP = 10, FL_DONE_FROM_DOWN = 11 };

__device__ __forceinline__ unsigned ld_flag(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long wall_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded by WALL-CLOCK time (%globaltimer), not by a poll count: a neighbour may legitimately lag by seconds
// (first-step graph instantiation, module load, host jitter).  After `timeout_ns` without progress -- or as soon as
// an earlier wait of this rank has timed out -- the error flag is raised (f2d_sync reports and clears it) and the
// kernel moves on instead of hanging the device.
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned e, unsigned* err, unsigned long long timeout_ns) {
    if ((int)(ld_flag(p) - e) >= 0) return;
    const unsigned long long t0 = wall_ns();
    unsigned it = 0;
    while ((int)(ld_flag(p) - e) < 0) {
        __nanosleep(64);
        if ((++it & 255u) == 0u && (wall_ns() - t0 > timeout_ns || ld_flag(err) != 0u)) {
            atomicExch(err, 1u);
            break;
        }
    }
}
}  // namespace

__global__ void __launch_bounds__(256) k_halo_xchg(XchgParams P) {
    __shared__ unsigned e_s;
    unsigned* mf = P.my_flags;
    if (threadIdx.x == 0) {
        const unsigned e = ld_flag(mf + FL_EPOCH) + 1u;
        e_s = e;
        if (blockIdx.x == 0) {
            if (P.up_flags) st_flag(P.up_flags + FL_READY_FROM_DOWN, e);
            if (P.down_flags) st_flag(P.down_flags + FL_READY_FROM_UP, e);
        }
        if (P.up_flags) wait_flag(mf + FL_READY_FROM_UP, e, mf + FL_ERROR, P.timeout_ns);
        if (P.down_flags) wait_flag(mf + FL_READY_FROM_DOWN, e, mf + FL_ERROR, P.timeout_ns);
    }
    __syncthreads();
    const unsigned e = e_s;
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // the push: local loads, 16-byte peer stores over NVLink.  Four independent loads are in flight per thread before
    // the first store (the copy is latency-, not bandwidth-bound: a few MB per exchange)
    for (int s = 0; s < P.nseg; ++s) {
        const float4* __restrict__ src = P.seg[s].src;
        float4* dst = P.seg[s].dst;
        const unsigned n4 = P.seg[s].n4;
        unsigned i = tid;
        for (; i + 3u * nth < n4; i += 4u * nth) {
            const float4 a = __ldg(src + i), b = __ldg(src + i + nth), c = __ldg(src + i + 2u * nth), d = __ldg(src + i + 3u * nth);
            dst[i] = a;
            dst[i + nth] = b;
            dst[i + 2u * nth] = c;
            dst[i + 3u * nth] = d;
        }
        for (; i < n4; i += nth) dst[i] = __ldg(src + i);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(mf + FL_BLOCKS, 1u);
        if (old == gridDim.x - 1) {  // every block of this rank has pushed and fenced
            st_flag(mf + FL_BLOCKS, 0u);
            __threadfence_system();
            if (P.up_flags) st_flag(P.up_flags + FL_DONE_FROM_DOWN, e);
            if (P.down_flags) st_flag(P.down_flags + FL_DONE_FROM_UP, e);
            if (P.up_flags) wait_flag(mf + FL_DONE_FROM_UP, e, mf + FL_ERROR, P.timeout_ns);
            if (P.down_flags) wait_flag(mf + FL_DONE_FROM_DOWN, e, mf + FL_ERROR, P.timeout_ns);
            __threadfence_system();
            st_flag(mf + FL_EPOCH, e);
        }
    }
}

void launch_halo_xchg(const XchgParams& p, cudaStream_t st) { k_halo_xchg<<<128, 256, 0, st>>>(p); }

}  // namespace f2d
