// One-shot all-reduce of the packed batch-norm statistics over NVLink peer memory (sm_100a, NVSwitch: every peer at full
// bandwidth), replacing one NCCL all-reduce per BN layer per direction (sync_batchnorm/batchnorm.py:90-125: the
// reference's reduce + broadcast through a Python thread rendezvous).  The messages are 2*Cp doubles (forward) or floats
// (backward) - at most 16 KB - so the collective is pure latency: NCCL spends 10-20 us per call inside the captured
// training graph, ~100 calls per iteration.
//
// Every rank owns one SYMMETRIC buffer (torch.distributed._symmetric_memory: the same allocation mapped into every
// peer's address space): data[2 parities][world source ranks][MAX_N doubles] + flag[2][world] (uint64 sequence numbers).
// One CTA per rank:
//   1. PUSH: write the local vector into slot [parity][my rank] of EVERY rank's buffer (plain stores through the peer
//      mapping = NVLink writes), __threadfence_system, then a release store of the sequence number into
//      flag[parity][my rank] of every rank;
//   2. WAIT: spin (acquire loads, system scope) on the `world` flags of the local buffer until they carry this sequence;
//   3. SUM the `world` slots in RANK ORDER - every rank adds the same numbers in the same order, so the result is
//      bit-identical across ranks (the N-rank == 1-rank invariant then only sees the reassociation of the partial sums).
// The sequence counter lives in device memory and is advanced by the kernel itself: the launch is CUDA-graph
// capturable, no host value changes between replays.  Two parities make the buffer safe for back-to-back calls: a rank
// can only start call s+1 after all ranks posted their flags of call s, i.e. after everybody finished reading call s-1.
#include "common.cuh"
#include "../../include/monkey_b200.h"

namespace {
constexpr int P2P_MAX_N = 4096;      // doubles per slot: 2 * Cp <= 2 * 1036 in every shipped configuration
constexpr int P2P_MAX_WORLD = 16;

struct P2PP {
    unsigned char* peers[P2P_MAX_WORLD];
    int rank, world, n, is_double;
    void* local;
    unsigned long long* seq;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k_stats_allreduce(const P2PP p) {
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(p.seq) + 1;
    const int par = (int)(seq & 1);
    const size_t slot_bytes = (size_t)P2P_MAX_N * sizeof(double);
    const size_t flag_off = 2 * (size_t)P2P_MAX_WORLD * slot_bytes;
    // 1. push
    for (int dst = 0; dst < p.world; ++dst) {
        double* slot = reinterpret_cast<double*>(p.peers[dst] + ((size_t)par * P2P_MAX_WORLD + p.rank) * slot_bytes);
        if (p.is_double) {
            const double* src = reinterpret_cast<const double*>(p.local);
            for (int i = threadIdx.x; i < p.n; i += blockDim.x) slot[i] = src[i];
        } else {
            const float* src = reinterpret_cast<const float*>(p.local);
            for (int i = threadIdx.x; i < p.n; i += blockDim.x) slot[i] = (double)src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < p.world) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(p.peers[threadIdx.x] + flag_off) + par * P2P_MAX_WORLD + p.rank;
        st_release_sys(f, seq);
        // 2. wait for source rank threadIdx.x
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(p.peers[p.rank] + flag_off) +
                                         par * P2P_MAX_WORLD + threadIdx.x;
        unsigned long long spins = 0;
        while (ld_acquire_sys(mine) < seq) {
            if (++spins > (1ull << 25)) __trap();   // ~20 s: a peer that never arrives must abort, not hang the box
        }
    }
    __syncthreads();
    // 3. rank-ordered sum
    const double* base = reinterpret_cast<const double*>(p.peers[p.rank] + (size_t)par * P2P_MAX_WORLD * slot_bytes);
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < p.world; ++r) s += base[(size_t)r * P2P_MAX_N + i];
        if (p.is_double) reinterpret_cast<double*>(p.local)[i] = s;
        else reinterpret_cast<float*>(p.local)[i] = (float)s;
    }
    if (threadIdx.x == 0) *p.seq = seq;
}
}  // namespace

// bytes every rank must allocate (symmetric) for mk_stats_allreduce
MK_EXPORT int mk_stats_allreduce_bytes(void) {
    return (int)(2 * (size_t)P2P_MAX_WORLD * P2P_MAX_N * sizeof(double) + 2 * P2P_MAX_WORLD * sizeof(unsigned long long));
}

// In-place sum over ranks of `n` doubles (is_double != 0) or floats at `local`.  peers = HOST array of `world` device
// pointers: entry r is rank r's symmetric buffer as mapped in THIS process (entry `rank` = the local one), zeroed once
// before the first call.  seq = device uint64, zero-initialised, owned by this communicator.
MK_EXPORT int mk_stats_allreduce(void* local, int n, int is_double, const unsigned long long* peers, int rank, int world,
                                 unsigned long long* seq, void* stream) {
    MK_REQUIRE(world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "mk_stats_allreduce: bad rank/world");
    MK_REQUIRE(n >= 0 && n <= P2P_MAX_N, "mk_stats_allreduce: n = %d exceeds %d", n, P2P_MAX_N);
    if (n == 0 || world == 1) return 0;
    P2PP p;
    for (int r = 0; r < world; ++r) p.peers[r] = reinterpret_cast<unsigned char*>(peers[r]);
    p.rank = rank; p.world = world; p.n = n; p.is_double = is_double; p.local = local; p.seq = seq;
    k_stats_allreduce<<<1, 256, 0, (cudaStream_t)stream>>>(p);
    return mk_check_launch("mk_stats_allreduce");
}
