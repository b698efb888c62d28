// The ONE collective of the path — the sum over the ranks of the per-rank Chamfer partials (a handful of
// floats) — as this library's own kernel over NVLink peer memory instead of an ncclAllReduce launch.
//
// Every rank owns a small "mailbox" buffer in device memory that all peers of the node map through CUDA IPC
// (dispu_peer_alloc / dispu_peer_open; the handles travel over torch.distributed once, outside the timed path).
// A step's all-reduce is then ONE tiny kernel per rank (one warp):
//   1. lane r writes this rank's n values and the step's epoch number into slot `rank` of peer r's mailbox — plain
//      stores over NVLink / NVSwitch, the payload first, a system-scope fence, then the flag;
//   2. lane r spins (ld.acquire.sys) until slot r of the OWN mailbox carries the epoch, i.e. peer r's values have landed;
//   3. the values are summed in rank order (the same order on every rank: bit-identical results everywhere, run to run).
// One-shot all-gather + local reduction: a single NVLink store latency, no ring, no second kernel.  An 8-byte payload
// is latency-only on any transport; what this removes is NCCL's launch + protocol floor (~25-30 us on top of the step at
// 8 ranks in round 1/2 measurements).  The mailbox is double-buffered by epoch parity: a fast rank's step t+1 values land in
// the other half, and it cannot reach t+2 before every peer has finished reading t (it needs their t+1 flags first).
// The epoch counter lives in device memory and is advanced by the kernel, so the launch is CUDA-graph replayable.
// A rank whose peer never arrives gives up after ~2 s and returns NaN instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace dispu {

constexpr int kPrMaxRanks = 32;
constexpr int kPrMaxVals = 6;
constexpr int kPrSlotFloats = 8;                        // 6 values | flag | pad = 32 bytes
constexpr size_t kPrBytes = 2 * kPrMaxRanks * kPrSlotFloats * sizeof(float);  // two parities

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(32)
peer_allreduce_kernel(int world, int rank, float *const *__restrict__ peers /* device array: mailbox of every rank */,
                      unsigned *__restrict__ epoch_ctr, const float *__restrict__ in, float *__restrict__ out, int n) {
    const int lane = threadIdx.x;
    const unsigned epoch = *epoch_ctr + 1u;  // (read by every lane before lane 0 advances it below)
    const int par = (int)(epoch & 1u);
    __syncwarp();
    if (lane < world) {
        float *slot = peers[lane] + ((size_t)par * kPrMaxRanks + rank) * kPrSlotFloats;  // my slot in peer `lane`'s mailbox
        for (int i = 0; i < n; ++i) slot[i] = in[i];
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned *>(slot + kPrMaxVals), epoch);
    }
    float v[kPrMaxVals];
    bool ok = true;
    if (lane < world) {
        const float *slot = peers[rank] + ((size_t)par * kPrMaxRanks + lane) * kPrSlotFloats;  // peer `lane`'s slot in MY mailbox
        const long long t0 = clock64();
        while (ld_acquire_sys(reinterpret_cast<const unsigned *>(slot + kPrMaxVals)) != epoch) {
            if (clock64() - t0 > 4000000000LL) { ok = false; break; }  // ~2 s: a peer died; do not hang the device
        }
        for (int i = 0; i < kPrMaxVals; ++i) v[i] = i < n ? __ldcv(slot + i) : 0.f;
    } else {
        for (int i = 0; i < kPrMaxVals; ++i) v[i] = 0.f;
    }
    ok = __all_sync(kFull, ok);
    // fixed-order sum over the ranks: lane 0 adds rank 0, 1, 2, ... (deterministic and identical on every rank)
    for (int i = 0; i < n; ++i) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc = __fadd_rn(acc, __shfl_sync(kFull, v[i], r));
        if (lane == 0) out[i] = ok ? acc : __int_as_float(0x7fc00000);
    }
    if (lane == 0) *epoch_ctr = epoch;
}

}  // namespace dispu

using namespace dispu;

extern "C" size_t dispu_peer_mailbox_bytes(void) { return kPrBytes; }

// Allocates this rank's mailbox (device memory, zeroed) and returns its CUDA IPC handle (64 bytes) for the peers.
extern "C" int dispu_peer_alloc(void **mailbox, unsigned char handle_out[64]) {
    DISPU_REQUIRE(mailbox && handle_out, DISPU_ENULL);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, kPrBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, kPrBytes);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle_out, &h, 64);
    *mailbox = p;
    return DISPU_OK;
}

// Maps a peer's mailbox (its handle from dispu_peer_alloc, received over the host-side process group) into this process.
extern "C" int dispu_peer_open(const unsigned char handle[64], void **mailbox) {
    DISPU_REQUIRE(handle && mailbox, DISPU_ENULL);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(mailbox, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? DISPU_OK : (int)e;
}

extern "C" int dispu_peer_close(void *mailbox) {
    DISPU_REQUIRE(mailbox, DISPU_ENULL);
    const cudaError_t e = cudaIpcCloseMemHandle(mailbox);
    return e == cudaSuccess ? DISPU_OK : (int)e;
}

extern "C" int dispu_peer_free(void *mailbox) {
    DISPU_REQUIRE(mailbox, DISPU_ENULL);
    const cudaError_t e = cudaFree(mailbox);
    return e == cudaSuccess ? DISPU_OK : (int)e;
}

// out[0..n) = sum over the ranks of in[0..n), in rank order.  peers: DEVICE array of `world` mailbox pointers (entry
// `rank` = the own mailbox); epoch_ctr: one device uint32, zero before the first call, owned by this rank.
extern "C" int dispu_peer_allreduce_sum(int world, int rank, void *const *peers, unsigned *epoch_ctr, const float *in,
                                        float *out, int n, dispu_stream_t stream) {
    DISPU_REQUIRE(world >= 1 && rank >= 0 && rank < world && n >= 1, DISPU_EINVAL);
    DISPU_REQUIRE(world <= kPrMaxRanks && n <= kPrMaxVals, DISPU_EUNSUPPORTED);
    DISPU_REQUIRE(peers && epoch_ctr && in && out, DISPU_ENULL);
    count_launch();
    peer_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(world, rank, reinterpret_cast<float *const *>(peers), epoch_ctr, in, out, n);
    return finish_launch();
}
