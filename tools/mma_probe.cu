// Development probe: cycles per tcgen05.mma kind::tf32 dispatch on B200 as a function of the form (A from TMEM / from
// shared memory), N, the accumulator pattern and the number of CTAs issuing.  One thread per CTA issues R MMAs back to
// back on zeroed operands, commits, waits; clock64 around the lot.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o gpurun_out/mma_probe tools/mma_probe.cu ; run on the box.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dispu_b200/csrc/tc_common.cuh"

using namespace dispu::tc;

struct Variant { const char *name; int ts; int n; int nacc; int pattern; };

__global__ void __launch_bounds__(128, 1) probe(int ts, int n, int nacc, int pattern, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (tid < 32) tmem_alloc(&s_tmem, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, n);
        const uint32_t sb = smem_u32(smem);
        const uint32_t lboB = (uint32_t)(n / 8) * 128u, lboA = 16u * 128u;
        const uint32_t a_smem = sb + 128 * 1024;  // A operand for the SS form: 128 rows x 32 K
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const int ks = r & 3;  // 4 K-steps of 8 inside a 32-wide chunk, as the dense kernel walks them
            const uint32_t d = tmem + (uint32_t)(nacc > 1 ? (r % nacc) * n : 0);
            const uint64_t db = make_smem_desc(sb + ((pattern & 1) ? ((r >> 2) & 3) * 32768u : 0u) + ks * 2 * lboB, lboB, 128u);
            if (ts) {
                const uint32_t ta = tmem + 256u + (uint32_t)((pattern & 2) ? (r & 7) * 8 : ks * 8);
                mma_tf32_ts(d, ta, db, idesc, r >= nacc);
            } else {
                const uint64_t da = make_smem_desc(a_smem + ks * 2 * lboA, lboA, 128u);
                mma_tf32_ss(d, da, db, idesc, r >= nacc);
            }
        }
        mma_commit(&s_bar);
        mbar_wait(&s_bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

int main() {
    const Variant vs[] = {
        {"TS N=256 one acc", 1, 256, 1, 0},   {"TS N=128 one acc", 1, 128, 1, 0}, {"TS N=128 two acc", 1, 128, 2, 0},
        {"TS N=64 one acc", 1, 64, 1, 0},     {"TS N=64 four acc", 1, 64, 4, 0},  {"TS N=32 one acc", 1, 32, 1, 0},
        {"TS N=32 four acc", 1, 32, 4, 0},    {"TS N=16 one acc", 1, 16, 1, 0},   {"SS N=256 one acc", 0, 256, 1, 0},
        {"SS N=128 one acc", 0, 128, 1, 0},   {"SS N=128 two acc", 0, 128, 2, 0}, {"SS N=64 one acc", 0, 64, 1, 0},
        {"SS N=32 one acc", 0, 32, 1, 0},     {"TS N=128 B walks 4 slots", 1, 128, 1, 1},
        {"TS N=128 A walks 64 cols", 1, 128, 1, 2}, {"TS N=256 B walks, A walks", 1, 256, 1, 3},
        {"TS N=192 one acc", 1, 192, 1, 0},   {"TS N=96 one acc", 1, 96, 1, 0},
    };
    long long *d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int reps = 4096;
    for (int grid : {1, 148}) {
        for (const Variant &v : vs) {
            std::vector<long long> h(148, 0);
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<grid, 128, 160 * 1024>>>(v.ts, v.n, v.nacc, v.pattern, reps, d_out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(cudaGetLastError())); return 1; }
            }
            cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0, mn = 1LL << 60;
            for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
            printf("grid %3d  %-28s  %7.1f .. %7.1f clk / MMA   (floor N/2 = %d)\n", grid, v.name, (double)mn / reps, (double)mx / reps, v.n / 2);
        }
    }
    return 0;
}
