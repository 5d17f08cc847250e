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
__global__ void __launch_bounds__(256) bulk_ring(const char* base, size_t bytes_per_cta, int stage_bytes, int S, int split, double* sink, int rot = 1, int mode = 0, double* outbuf = nullptr, unsigned long long* flag = nullptr)
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
        if (mode & 1) {       // ~350 cycles of dependent arithmetic per step
            double z = acc;
            for (int it = 0; it < 40; ++it) z = z * 1.0000001 + 1e-9;
            acc = z;
        }
        if (mode & 2) {       // one 16-byte streaming store per thread and step (4 KB per step, contiguous)
            double* o = outbuf + ((size_t)blockIdx.x * 4096 + (size_t)(s & 4095)) * 512 + t * 2;
            __stcg(reinterpret_cast<double2*>(o), make_double2(acc, acc));
        }
        __syncthreads();
        if ((mode & 4) && t == 0 && (s & 7) == 7) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flag + blockIdx.x), "l"((unsigned long long)s) : "memory");
        if (t == ((s + S) % rot) * 32 && s + S < nst) issue(s + S);
    }
    if (acc == 1.2345) sink[0] = acc;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// producer warp (warp 8) + 256 consumer threads, full/empty mbarriers per slot
__global__ void __launch_bounds__(288) bulk_pc(const char* base, size_t bytes_per_cta, int stage_bytes, int S, int delay, double* sink, long long* stamps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;
    const int t = threadIdx.x;
    const char* src = base + (size_t)blockIdx.x * bytes_per_cta;
    const int nst = (int)(bytes_per_cta / stage_bytes);
    if (t == 0) {
        for (int q = 0; q < S; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t >= 256) {
        if (t == 256) {
            for (int s = 0; s < nst; ++s) {
                const int q = s % S;
                if (s >= S) mbar_wait(&empty[q], ((s / S) - 1) & 1);
                const long long c0 = clock64();
                mbar_expect_tx(&full[q], stage_bytes);
                bulk_g2s(smem + (size_t)q * stage_bytes, src + (size_t)s * stage_bytes, stage_bytes, &full[q]);
                const long long c1 = clock64();
                if (stamps && blockIdx.x == 0 && s < 64) { stamps[2 * s] = c0; stamps[2 * s + 1] = c1; }
            }
        }
        return;
    }
    double acc = 0.0;
    for (int s = 0; s < nst; ++s) {
        const int q = s % S;
        const long long k0 = clock64();
        mbar_wait(&full[q], (s / S) & 1);
        const long long k1 = clock64();
        acc += reinterpret_cast<const double*>(smem + (size_t)q * stage_bytes)[t];
        double z = acc;
        for (int it = 0; it < delay; ++it) z = z * 1.0000001 + 1e-9;
        acc = z;
        const long long k2 = clock64();
        __syncwarp();
        if ((t & 31) == 0) mbar_arrive(&empty[q]);
        const long long k3 = clock64();
        if (stamps && blockIdx.x == 0 && t == 0 && s >= 32 && s < 40) { stamps[128 + 4 * (s - 32)] = k0; stamps[129 + 4 * (s - 32)] = k1; stamps[130 + 4 * (s - 32)] = k2; stamps[131 + 4 * (s - 32)] = k3; }
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

__global__ void flush(double* p, size_t bytes)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes / 8; i += (size_t)gridDim.x * blockDim.x) p[i] = 1.0;
}

int main()
{
    const size_t per_cta = 128ull << 20;      // region stride per CTA
    const int maxcta = 148;
    char* buf;
    double* sink;
    cudaMalloc(&buf, per_cta * maxcta);
    cudaMalloc(&sink, 8);
    double* outbuf; unsigned long long* flags;
    cudaMalloc(&outbuf, (size_t)148 * 4096 * 512 * 8);
    cudaMalloc(&flags, 148 * 8);
    double* flushbuf;
    cudaMalloc(&flushbuf, 256u << 20);
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
    const int ctas_list[] = {1};
    for (int ctas : ctas_list) {
        // every timed pass streams >= 512 MiB that the (different) warm-up pass did not touch: HBM-cold, L2 is 126 MB
        const size_t bytes = ctas <= 4 ? per_cta : (ctas <= 16 ? per_cta / 2 : per_cta / 8);
        const size_t warm = 1 << 20;
        struct Cfg { int stage, S, split, rot, mode; } cfgs[] = {{28672, 7, 1, 1, 0}, {28672, 7, 1, 1, 1}};
        for (auto c : cfgs) {
            
            const size_t smem = (size_t)c.stage * c.S + 8 * c.S + 128;
            cudaFuncSetAttribute(bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const size_t use = bytes / c.stage * c.stage;
            bulk_ring<<<ctas, 256, smem>>>(buf + per_cta / 2 + (64 << 20) - (64 << 20), warm / c.stage * c.stage, c.stage, c.S, c.split, sink, c.rot);
            flush<<<1184, 256>>>(flushbuf, 256u << 20);
            cudaEventRecord(a);
            bulk_ring<<<ctas, 256, smem>>>(buf, use, c.stage, c.S, c.split, sink, c.rot, c.mode, outbuf, flags);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            char nm[64]; snprintf(nm, 64, "bulk stage %6d S %2d mode %d", c.stage, c.S, c.mode);
            report(nm, ctas, ms, use);
        }
        {
            struct PC { int stage, S, delay; } pcs[] = {{28672, 7, 0}, {28672, 7, 40}, {28672, 7, 120}, {28672, 2, 40}, {86016, 2, 0}, {86016, 2, 120}, {86016, 2, 360}, {57344, 3, 80}, {57344, 3, 240}};
            long long* stamps; cudaMalloc(&stamps, 256 * 8); cudaMemset(stamps, 0, 256 * 8);
            for (auto c : pcs) {
                const size_t smem = (size_t)c.stage * c.S + 16 * c.S + 128;
                cudaFuncSetAttribute(bulk_pc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                const size_t use = bytes / c.stage * c.stage;
                flush<<<1184, 256>>>(flushbuf, 256u << 20);
                cudaEventRecord(a);
                bulk_pc<<<ctas, 288, smem>>>(buf, use, c.stage, c.S, c.delay, sink, stamps);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                char nm[64]; snprintf(nm, 64, "prodcons stage %6d S %d delay %3d", c.stage, c.S, c.delay);
                report(nm, ctas, ms, use);
                if (ctas == 1) {
                    long long h[128]; cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("   issue stamps (start delta, issue duration):");
                    for (int q = 1; q < 12; ++q) printf(" (%lld,%lld)", h[2 * q] - h[2 * q - 2], h[2 * q + 1] - h[2 * q]);
                    printf("\n");
                    long long k[32]; cudaMemcpy(k, stamps + 128, sizeof(k), cudaMemcpyDeviceToHost);
                    printf("   consumer (wait, work, arrive, gap):");
                    for (int q = 0; q < 7; ++q) printf(" (%lld,%lld,%lld,%lld)", k[4 * q + 1] - k[4 * q], k[4 * q + 2] - k[4 * q + 1], k[4 * q + 3] - k[4 * q + 2], k[4 * q + 4] - k[4 * q + 3]);
                    printf("\n");
                }
            }
        }
        {
            ldg_stream<8><<<ctas, 256>>>(buf, warm, sink);
            flush<<<1184, 256>>>(flushbuf, 256u << 20);
            cudaEventRecord(a);
            ldg_stream<8><<<ctas, 256>>>(buf, bytes, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            report("ldg.128 x8 per thread", ctas, ms, bytes);
            ldg_stream<16><<<ctas, 256>>>(buf, warm, sink);
            flush<<<1184, 256>>>(flushbuf, 256u << 20);
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
