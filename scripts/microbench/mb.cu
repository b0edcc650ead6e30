// Micro-benchmarks that steer the kernel design (development only; results under profiles/microbench_r02.json):
//   1. issue rate of scalar FADD / FFMA vs the packed FADD2 / FFMA2 (add.rn.f32x2 / fma.rn.f32x2) per SM and clock
//   2. streaming copy global -> shared -> global with persistent CTAs: classic LDG/STS staging vs cp.async.bulk (TMA 1-d bulk
//      copy) + mbarrier double buffering, for the x-line tile size of the c3 plan (8 lines x 4352 B)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb mb.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) alu_kernel(float* out, int iters) {
    float a[16], b = 1.0001f + threadIdx.x * 1e-7f, c = 0.9999f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], b);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], c, b);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long x, y;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(b));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(x), "l"(y));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long x, y, z;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(b));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c), "f"(c));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(z), "l"(y));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
            }
        } else if (MODE == 4) {      // mixed: 8 FADD + 8 IADD-like (alu pipe) to see dual issue across pipes
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fadd_rn(a[i], b);
#pragma unroll
            for (int i = 8; i < 16; ++i) a[i] = __int_as_float(__float_as_int(a[i]) + 3);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// persistent copy, tile = TILE bytes; MODE 0: LDG.128 -> STS -> sync -> LDS -> STG (classic), MODE 1: bulk copy double buffered -> LDS -> STG
template <int MODE, int THREADS, int TILE, int STAGES>
__global__ void __launch_bounds__(THREADS) copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long ntiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int Q = TILE / 16;    // float4 per tile
    __shared__ uint64_t bars[STAGES];
    const int tid = threadIdx.x;
    if (MODE == 1) {
        if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncthreads();
        if (tid == 0) {
            for (int s = 0; s < STAGES; ++s) {
                const long long t = blockIdx.x + (long long)s * gridDim.x;
                if (t < ntiles) { mbar_expect_tx(&bars[s], TILE); bulk_g2s(smem + (size_t)s * TILE, in + t * Q, TILE, &bars[s]); }
            }
        }
    }
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        float4* buf = reinterpret_cast<float4*>(smem + (size_t)s * TILE);
        if (MODE == 0) {
            for (int i = tid; i < Q; i += THREADS) buf[i] = __ldg(in + t * Q + i);
            __syncthreads();
        } else {
            mbar_wait(&bars[s], (it / STAGES) & 1);
        }
        for (int i = tid; i < Q; i += THREADS) { float4 v = buf[(i * 7 + 3) % Q]; v.x += 1.f; out[t * Q + i] = v; }
        __syncthreads();
        if (MODE == 1 && tid == 0) {
            const long long tn = t + (long long)STAGES * gridDim.x;
            if (tn < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[s], TILE); bulk_g2s(smem + (size_t)s * TILE, in + tn * Q, TILE, &bars[s]);
            }
        }
    }
}

template <int MODE, int THREADS, int TILE, int STAGES>
void run_copy(const char* name, const float4* in, float4* out, size_t bytes, int ctas_per_sm) {
    const long long ntiles = bytes / TILE;
    const size_t smem = (size_t)TILE * STAGES;
    CK(cudaFuncSetAttribute(copy_kernel<MODE, THREADS, TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        copy_kernel<MODE, THREADS, TILE, STAGES><<<148 * ctas_per_sm, THREADS, smem>>>(in, out, ntiles);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"bench\": \"copy\", \"variant\": \"%s\", \"threads\": %d, \"tile\": %d, \"stages\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBs\": %.1f}\n",
           name, THREADS, TILE, STAGES, ctas_per_sm, ms, 2.0 * (double)ntiles * TILE / ms / 1e6);
}

int main() {
    float* out; CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(float)));
    const int iters = 4096;
    const char* names[] = {"FADD", "FFMA", "FADD2(f32x2)", "FFMA2(f32x2)", "FADD+IADD mix"};
    for (int mode = 0; mode < 5; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: alu_kernel<0><<<148 * 8, 256>>>(out, iters); break;
                case 1: alu_kernel<1><<<148 * 8, 256>>>(out, iters); break;
                case 2: alu_kernel<2><<<148 * 8, 256>>>(out, iters); break;
                case 3: alu_kernel<3><<<148 * 8, 256>>>(out, iters); break;
                case 4: alu_kernel<4><<<148 * 8, 256>>>(out, iters); break;
            }
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double lane_ops = 148.0 * 8 * 256 * (double)iters * 16;     // scalar fp32 results produced
        int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("{\"bench\": \"alu\", \"op\": \"%s\", \"ms\": %.4f, \"fp32_results_per_clk_per_sm\": %.1f, \"clock_khz_nominal\": %d}\n", names[mode], ms,
               lane_ops / (ms * 1e-3) / (clk * 1e3) / 148.0, clk);
    }
    const size_t bytes = (size_t)2 << 30;
    float4 *in, *o2; CK(cudaMalloc(&in, bytes)); CK(cudaMalloc(&o2, bytes));
    CK(cudaMemset(in, 0, bytes));
    run_copy<0, 288, 34816, 1>("ldg-sts", in, o2, bytes, 3);
    run_copy<0, 288, 34816, 1>("ldg-sts", in, o2, bytes, 4);
    run_copy<1, 288, 34816, 2>("tma-bulk", in, o2, bytes, 1);
    run_copy<1, 288, 34816, 2>("tma-bulk", in, o2, bytes, 2);
    run_copy<1, 288, 34816, 2>("tma-bulk", in, o2, bytes, 3);
    run_copy<1, 288, 34816, 3>("tma-bulk", in, o2, bytes, 2);
    run_copy<1, 576, 34816, 3>("tma-bulk", in, o2, bytes, 1);
    run_copy<1, 576, 34816, 4>("tma-bulk", in, o2, bytes, 1);
    run_copy<1, 288, 17408, 3>("tma-bulk", in, o2, bytes, 3);
    run_copy<1, 288, 17408, 4>("tma-bulk", in, o2, bytes, 4);
    return 0;
}
