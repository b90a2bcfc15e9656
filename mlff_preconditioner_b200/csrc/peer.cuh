// NVLink peer-memory collectives for the solver's inner loops (one process per GPU, CUDA IPC).
//
// Every rank owns one communication buffer; all ranks map all buffers (cudaIpcOpenMemHandle), so a kernel can store
// into a peer's buffer or load from it directly over NVLink / NVSwitch.  The collectives of the PCG iteration and of
// the pivot loop are a few dozen bytes to a few hundred kilobytes -- latency, not bandwidth -- so instead of a library
// call per collective (launch + protocol, ~10-20 us each, 5 per CG iteration) the PRODUCING kernel pushes its result
// into the peers' slots and raises a flag, and the CONSUMING kernel waits for the flags in its prologue and combines
// the slots in rank order (deterministic, bit-identical on every rank):
//   scalars   dot kernels push (rho, ||r||^2) / (p.q)        -> p-update / x,r-update kernels combine them
//   k-vector  T r of the preconditioner apply: one small kernel pushes, waits and sums in place
//   p         the p-update kernel stores its slice of the search direction into every rank's replicated vector
//   K p       symmetric tile operator: the finish kernel pulls the peers' partial products of its rows (a fused
//             reduce-scatter) and applies alpha / shift
//   pivots    the prepare kernel pushes {candidate, factor row}; the update kernel reads the gathered messages
// Flags are monotonically increasing 64-bit epochs (one per channel and source rank): a system-scope release fence
// followed by relaxed flag stores on the producer, relaxed polls followed by an acquire fence on the consumer.  Without a peer mapping (single GPU, IPC refused) the NCCL path is used.
#pragma once
#include <stdint.h>

namespace mlffpc {

constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_CH_SCAL_A = 0;   // (rho, ||r||^2)
constexpr int PEER_CH_SCAL_B = 1;   // p.q
constexpr int PEER_CH_KVEC = 2;
constexpr int PEER_CH_P = 3;        // search direction stored
constexpr int PEER_CH_YP = 4;       // partial products ready
constexpr int PEER_CH_MSG = 5;      // pivot message
constexpr int PEER_NCH = 8;
constexpr int PEER_MSG_DOUBLES = 80;  // >= LA_MSG of pchol.cu

struct PeerLayout {
    int64_t off_flags, off_counter, off_scal, off_kvec, off_p, off_yp, off_msg, total;
    int64_t k_pad, n_full;   // doubles per k-vector slot; world * n_pad
};

// device-visible view: base pointers of every rank's buffer + this rank's identity + the layout
struct PeerView {
    char* base[PEER_MAX_RANKS];
    int rank, world;
    PeerLayout lay;
    __device__ __forceinline__ uint64_t* flags(int r, int ch) const {
        return reinterpret_cast<uint64_t*>(base[r] + lay.off_flags) + (int64_t)ch * PEER_MAX_RANKS;
    }
    __device__ __forceinline__ unsigned* counter(int which) const {
        return reinterpret_cast<unsigned*>(base[rank] + lay.off_counter) + which;
    }
    __device__ __forceinline__ double* scal(int r, int parity, int src) const {
        return reinterpret_cast<double*>(base[r] + lay.off_scal) + ((int64_t)parity * PEER_MAX_RANKS + src) * 8;
    }
    __device__ __forceinline__ double* kvec(int r, int parity, int src) const {
        return reinterpret_cast<double*>(base[r] + lay.off_kvec) + ((int64_t)parity * PEER_MAX_RANKS + src) * lay.k_pad;
    }
    __device__ __forceinline__ double* p_full(int r) const { return reinterpret_cast<double*>(base[r] + lay.off_p); }
    __device__ __forceinline__ double* yp(int r) const { return reinterpret_cast<double*>(base[r] + lay.off_yp); }
    __device__ __forceinline__ double* msg(int r, int parity, int src) const {
        return reinterpret_cast<double*>(base[r] + lay.off_msg) + ((int64_t)parity * PEER_MAX_RANKS + src) * PEER_MSG_DOUBLES;
    }
};

__device__ __forceinline__ uint64_t peer_ld_acquire(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st_release(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// loads that must not be served from this SM's L1 (the line may be from an earlier epoch or live on a peer)
__device__ __forceinline__ double peer_ld(const double* p) { return __ldcg(p); }

__device__ __forceinline__ void peer_st_relaxed(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t peer_ld_relaxed(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// raise this rank's flag of channel ch on every rank (call from ONE thread after the data stores of the whole grid /
// CTA are complete and fenced).  ONE release fence, then relaxed flag stores that travel together: a st.release per
// peer waits for the previous peer's store to be acknowledged over NVLink -- 8 ranks, 8 round trips, ~25 us per signal
// (measured: 42 us per pivot step on 8 GPUs against 10 us on one).
__device__ __forceinline__ void peer_signal_all(const PeerView& pv, int ch, uint64_t epoch) {
    __threadfence_system();
    for (int r = 0; r < pv.world; ++r) peer_st_relaxed(pv.flags(r, ch) + pv.rank, epoch);
}
// wait until every rank's flag of channel ch in MY buffer has reached epoch (one thread; follow with a barrier): the
// flags are polled together with relaxed loads, one acquire fence orders the data loads behind them
__device__ __forceinline__ void peer_wait_all(const PeerView& pv, int ch, uint64_t epoch) {
    const uint64_t* f = pv.flags(pv.rank, ch);
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int r = 0; r < PEER_MAX_RANKS; ++r)
            if (r < pv.world) ok &= (peer_ld_relaxed(f + r) >= epoch);
    } while (!ok);
    __threadfence_system();
}

// host side (peer.cu)
struct Peer {
    PeerView view;          // host copy (passed to kernels by value)
    void* local = nullptr;  // this rank's buffer (cudaMalloc)
    void* opened[PEER_MAX_RANKS] = {nullptr};
    int64_t k_max = 0;
    uint64_t epoch = 0;     // bumped by the host once per collective round; identical on all ranks
    uint64_t uses[PEER_NCH] = {0};  // per-channel round counter (slot parity)
    bool enabled = false;
};

}  // namespace mlffpc
