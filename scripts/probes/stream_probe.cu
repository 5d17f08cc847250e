// Developer microbenchmark (not product code): how fast can ONE SM stream a private contiguous region from HBM, as a function
// of the number of concurrently streaming SMs -- for cp.async.bulk rings (what ilu_sweep_kernel uses), plain LDG.128 and
// cp.async 16 B.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// ring of S stages, each stage fetched as `split` bulk copies
__global__ void __launch_bounds__(256) bulk_ring(const char* base, size_t bytes_per_cta, int stage_bytes, int S, int split, double* sink, int rot = 1)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    const int t = threadIdx.x;
    const char* src = base + (size_t)blockIdx.x * bytes_per_cta;
    const int nst = (int)(bytes_per_cta / stage_bytes);
    if (t == 0) {
        for (int q = 0; q < S; ++q) mbar_init(&mbar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int s) {
        const int q = s % S;
        mbar_expect_tx(&mbar[q], stage_bytes);
        const int part = stage_bytes / split;
        for (int c = 0; c < split; ++c) bulk_g2s(smem + (size_t)q * stage_bytes + c * part, src + (size_t)s * stage_bytes + c * part, part, &mbar[q]);
    };
    if (t == 0) for (int q = 0; q < S && q < nst; ++q) issue(q);
    double acc = 0.0;
    for (int s = 0; s < nst; ++s) {
        const int q = s % S;
        mbar_wait(&mbar[q], (s / S) & 1);
        acc += reinterpret_cast<const double*>(smem + (size_t)q * stage_bytes)[t];
        __syncthreads();
        if (t == ((s + S) % rot) * 32 && s + S < nst) issue(s + S);
    }
    if (acc == 1.2345) sink[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256) ldg_stream(const char* base, size_t bytes_per_cta, double* sink)
{
    const double2* src = reinterpret_cast<const double2*>(base + (size_t)blockIdx.x * bytes_per_cta);
    const size_t n = bytes_per_cta / 16;
    double acc = 0.0;
    for (size_t i = threadIdx.x; i + (size_t)(U - 1) * 256 < n; i += (size_t)U * 256) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(src + i + (size_t)u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
    }
    if (acc == 1.2345) sink[0] = acc;
}

int main()
{
    const size_t per_cta = 64ull << 20;       // 64 MiB per CTA
    const int maxcta = 296;
    char* buf;
    double* sink;
    cudaMalloc(&buf, per_cta * maxcta);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, per_cta * maxcta);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clk = 1.965e9;
    printf("clock attr %d kHz; B/clk computed at 1.965 GHz\n", clk_khz);
    auto report = [&](const char* name, int ctas, float ms, size_t bytes) {
        const double gbs = ctas * (double)bytes / ms / 1e6;
        printf("%-34s ctas %3d  %8.3f ms  total %7.0f GB/s  per-CTA %6.1f GB/s = %5.1f B/clk\n", name, ctas, ms, gbs, gbs / ctas, gbs / ctas * 1e9 / clk);
    };
    const int ctas_list[] = {1, 16, 74, 148};
    for (int ctas : ctas_list) {
        const size_t bytes = ctas <= 16 ? per_cta / 4 : per_cta / 8;
        struct Cfg { int stage, S, split, rot; } cfgs[] = {{28672, 7, 1, 1}, {28672, 7, 1, 7}, {8192, 24, 1, 8}, {8192, 24, 2, 1}, {16384, 12, 1, 4}, {32768, 6, 1, 1}, {49152, 4, 1, 1}, {65536, 3, 1, 1}, {98304, 2, 1, 1}, {106496, 2, 1, 1}, {106496, 2, 2, 1}, {106496, 2, 4, 1}};
        for (auto c : cfgs) {
            
            const size_t smem = (size_t)c.stage * c.S + 8 * c.S + 128;
            cudaFuncSetAttribute(bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const size_t use = bytes / c.stage * c.stage;
            bulk_ring<<<ctas, 256, smem>>>(buf, use, c.stage, c.S, c.split, sink, c.rot);
            cudaEventRecord(a);
            bulk_ring<<<ctas, 256, smem>>>(buf, use, c.stage, c.S, c.split, sink, c.rot);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            char nm[64]; snprintf(nm, 64, "bulk stage %6d S %2d split %2d rot %d", c.stage, c.S, c.split, c.rot);
            report(nm, ctas, ms, use);
        }
        {
            ldg_stream<8><<<ctas, 256>>>(buf, bytes, sink);
            cudaEventRecord(a);
            ldg_stream<8><<<ctas, 256>>>(buf, bytes, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            report("ldg.128 x8 per thread", ctas, ms, bytes);
            ldg_stream<16><<<ctas, 256>>>(buf, bytes, sink);
            cudaEventRecord(a);
            ldg_stream<16><<<ctas, 256>>>(buf, bytes, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            report("ldg.128 x16 per thread", ctas, ms, bytes);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    return 0;
}
