// dsc_shard.cuh -- ONE frame pair over several GPUs (second row): the correspondences are sharded by
// contiguous ranges of tiles of the internal (Morton) order, one process per GPU, every exchange fused into the
// producing / consuming kernels over NVLink peer memory -- no collective library call anywhere on the solve path.
//
// Every rank holds the whole pair (same upload, same deterministic set-up, hence the same internal numbering and the
// same sliced ELL) but computes only the rows [row_begin[rank], row_begin[rank + 1]).  What crosses NVLink:
//   * halo rows: a row whose neighbour lives on another rank is PUSHED by its owner into that rank's local buffer with
//     remote stores from inside the kernel that produces it (z of the PCG: cg_init / cg_update; trial state: apply_update),
//     so every consumer kernel (cg_spmv, cost, linearise) reads local memory only;
//   * partial sums: the rank's total (its block partials folded in a fixed order) goes into every rank's mailbox and
//     consumers add the G mailbox entries in rank order -> the same bits on every rank, every run.  Inside the PCG loop
//     (gamma from cg_init / cg_update; delta + the 8 global rows from cg_spmv) the LAST BLOCK of the producing kernel to
//     finish does that, so an iteration stays two launches; the per-trial sums (linearisation, trial scale, cost) take
//     a one-block exchange kernel (shard_allreduce_kernel), whose flag also publishes the halo rows its predecessor pushed;
//   * flags: a message is complete when its flag (a sequence number, release / acquire at system scope) has arrived;
//     consumer kernels wait for it in their prologue.  All ranks run the same kernel sequence (every decision is taken
//     from the same totals), so message n of a class on one rank pairs with message n of that class on every other.
// Mailbox slots and the z buffers are double-buffered by message parity: a producer can only be one message ahead of
// the slowest consumer (it needs that consumer's next message to proceed), so parity n is never overwritten while read.
// Waits are bounded (a rank that died must not hang the others): on a time-out the error word is set and the kernels run
// to completion on whatever they have; the host reports DSC_ERR_SHARD.
#pragma once
#include <cuda_runtime.h>

namespace dsc {

constexpr int kMaxShards = 8;
constexpr int kMboxDoubles = 48;             // doubles per message slot
enum ShardClass { SF_Z = 0 /* z pushed + gamma */, SF_S = 1 /* operator partials */, SF_P = 2 /* trial state pushed */,
                  SF_L = 3 /* linearisation partials */, SF_C = 4 /* cost + scale partials */, SF_A = 5 /* final all-gather */,
                  SF_COUNT = 6 };

struct ShardDev {
    int rank, world;
    int row_begin[kMaxShards + 1];               // multiples of the tile size, row_begin[world] = n
    double* Pbuf[2][kMaxShards];                 // the two state buffers of every rank (peer-mapped; [.][rank] is local)
    double* zbuf[2][kMaxShards];                 // PCG vector z by message parity
    double* mbox[kMaxShards];                    // [SF_COUNT][2][world][kMboxDoubles]
    unsigned long long* flags[kMaxShards];       // [SF_COUNT][world]
    unsigned long long* sent;                    // local [SF_COUNT]: messages this rank has sent per class
    unsigned int* ticket;                        // local [SF_COUNT]: blocks that have finished (last-block election)
    const unsigned char* exportmask;             // local [n]: bit q = rank q holds row i as a halo row
    int* error;                                  // local: set by a wait that timed out
    long long spin_limit;                        // clock64 ticks a wait may take
};

__device__ __forceinline__ unsigned long long shard_ld_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void shard_st_flag(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double* shard_slot(const ShardDev& S, int dst, int cls, unsigned long long seq, int src) {
    return S.mbox[dst] + (((size_t)cls * 2 + (size_t)(seq & 1ull)) * S.world + src) * kMboxDoubles;
}

// Block-wide: returns when message `seq` of class cls from EVERY rank has arrived here (or the wait timed out).
__device__ __forceinline__ void shard_wait(const ShardDev& S, int cls, unsigned long long seq) {
    if ((int)threadIdx.x < S.world) {
        const unsigned long long* f = S.flags[S.rank] + (size_t)cls * S.world + threadIdx.x;
        const long long t0 = clock64();
        while (shard_ld_flag(f) < seq) {
            if (*reinterpret_cast<volatile int*>(S.error)) break;                 // a wait has already timed out: do not wait again
            if (clock64() - t0 > S.spin_limit) { atomicExch(S.error, 1); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ unsigned long long shard_sent(const ShardDev& S, int cls) {
    return *reinterpret_cast<volatile unsigned long long*>(S.sent + cls);
}

// Last-block election of a producer kernel: every block calls this after its global writes (block partials, remote
// stores); exactly one block -- the last to arrive -- gets true, with all other blocks' writes visible to it.
// pushed: some thread of this block has stored into a PEER's memory (halo rows) -- its stores get a system-scope fence
// before the ticket; a block that only wrote local memory (its partials) needs the cheaper device-scope one.
__device__ __forceinline__ bool shard_last_block(const ShardDev& S, int cls, bool pushed = false) {
    __shared__ int last;
    const int any = __syncthreads_or(pushed ? 1 : 0);  // (also orders the block's writes before thread 0's fence: cumulativity)
    if (threadIdx.x == 0) {
        if (any) __threadfence_system(); else __threadfence();
        const unsigned int t = atomicAdd(S.ticket + cls, 1u);
        last = (t == gridDim.x - 1) ? 1 : 0;
        if (last) { S.ticket[cls] = 0; __threadfence(); }         // ready for the next launch; acquire side of the election
    }
    __syncthreads();
    return last != 0;
}

// Called by the elected block: message (vals[0 .. count), count <= kMboxDoubles, may be 0) to every rank, then the flag.
__device__ __forceinline__ void shard_send(const ShardDev& S, int cls, const double* vals, int count) {
    const unsigned long long seq = shard_sent(S, cls) + 1ull;
    for (int k = threadIdx.x; k < count * S.world; k += blockDim.x) {
        const int dst = k / count, e = k % count;
        shard_slot(S, dst, cls, seq, S.rank)[e] = vals[e];
    }
    // ONE system-scope release for the whole message: the barrier orders the block's stores (and, through the election,
    // every other block's) before the release stores of the flags -- a system-scope fence costs ~10 us on this machine,
    // and this block is the serial tail of the launch
    __syncthreads();
    if ((int)threadIdx.x < S.world) shard_st_flag(S.flags[threadIdx.x] + (size_t)cls * S.world + S.rank, seq);
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(S.sent + cls) = seq;
    __syncthreads();
}

// Sum over the ranks, in rank order, of entry e of message `seq` of class cls (after shard_wait): same bits everywhere.
__device__ __forceinline__ double shard_total(const ShardDev& S, int cls, unsigned long long seq, int e) {
    double s = 0.0;
    for (int r = 0; r < S.world; ++r) s += __ldcg(shard_slot(S, S.rank, cls, seq, r) + e);
    return s;
}
__device__ __forceinline__ double shard_max(const ShardDev& S, int cls, unsigned long long seq, int e) {
    double s = 0.0;
    for (int r = 0; r < S.world; ++r) s = fmax(s, __ldcg(shard_slot(S, S.rank, cls, seq, r) + e));
    return s;
}

// One-block exchange: the rank's total of `count` per-block partials part[nb][stride] (entry maxidx: maximum instead of
// sum) -> every rank's mailbox -> wait for all ranks -> out[count] = totals over the ranks (rank order).  Launched right
// after the kernel that wrote `part`; the flag it raises also tells the peers that the halo rows that kernel pushed into
// their buffers are complete (stream order + the system-scope fence of the pushing threads).
__global__ void __launch_bounds__(256)
shard_allreduce_kernel(const __grid_constant__ ShardDev S, int cls, const double* __restrict__ part, int nb, int stride, int count,
                       int maxidx, double* __restrict__ out) {
    __shared__ double vals[kMboxDoubles];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = warp; e < count; e += (int)(blockDim.x >> 5)) {        // one warp per entry, fixed order
        double v = 0.0;
        for (int i = lane; i < nb; i += 32) {
            const double x = part[(size_t)i * stride + e];
            v = e == maxidx ? fmax(v, x) : v + x;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, v, o);
            v = e == maxidx ? fmax(v, y) : v + y;
        }
        if (lane == 0) vals[e] = v;
    }
    __syncthreads();
    shard_send(S, cls, vals, count);
    const unsigned long long seq = shard_sent(S, cls);
    shard_wait(S, cls, seq);
    for (int e = threadIdx.x; e < count; e += blockDim.x) out[e] = e == maxidx ? shard_max(S, cls, seq, e) : shard_total(S, cls, seq, e);
}

// the ranks that hold row i as a halo row: bit q set <=> a neighbour of row i is owned by rank q != owner(i)
__global__ void __launch_bounds__(256)
shard_exportmask_kernel(const __grid_constant__ ShardDev S, int n, const int* __restrict__ sliceptr, const int* __restrict__ ecol,
                        unsigned char* __restrict__ mask) {
    const int r0 = S.row_begin[S.rank], r1 = S.row_begin[S.rank + 1];
    for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
        const int sl = i >> 5, lane = i & 31;
        unsigned m = 0;
        for (int bk = sliceptr[sl]; bk < sliceptr[sl + 1]; ++bk) {
            const int j = ecol[(size_t)bk * 32 + lane];
            if (j >= r0 && j < r1) continue;
            int q = 0;
            while (q + 1 < S.world && j >= S.row_begin[q + 1]) ++q;
            m |= 1u << q;
        }
        mask[i] = (unsigned char)m;
    }
}

// final all-gather of the state: every rank pushes its own rows of P (both planes) to all peers (shard_allreduce_kernel
// of the update norm follows and publishes them)
__global__ void __launch_bounds__(256)
shard_allgather_state_kernel(const __grid_constant__ ShardDev S, int n, int pidx) {
    const int r0 = S.row_begin[S.rank], r1 = S.row_begin[S.rank + 1];
    const double4* src = reinterpret_cast<const double4*>(S.Pbuf[pidx][S.rank]);
    for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
        const double4 a = src[i], b = src[(size_t)n + i];
        for (int q = 0; q < S.world; ++q) {
            if (q == S.rank) continue;
            double4* dst = reinterpret_cast<double4*>(S.Pbuf[pidx][q]);
            dst[i] = a; dst[(size_t)n + i] = b;
        }
    }
    __threadfence_system();
}

}  // namespace dsc
