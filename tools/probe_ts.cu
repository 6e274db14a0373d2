// probe_ts.cu -- hardware feature probe for the TMEM-fed kernel (gemm4.cu), one tiny MMA per variant, each checked on the host:
//   tcgen05.st / tcgen05.ld round trip, tcgen05.mma with A in tensor memory (.ts form), A = f16 with B = f16 / bf16 (mixed
//   operand types of kind::f16), cta_group::1.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_ts tools/probe_ts.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// variant: 0 = st/ld round trip only; 1 = TS mma f16 x f16; 2 = TS mma A f16 x B bf16; 3 = TS mma bf16 x bf16; 4 = SS sanity f16
__global__ void __launch_bounds__(128) probe(int variant, const uint16_t *A /*[128][16]*/, const uint16_t *B /*[32][16]*/, float *D /*[128][32]*/,
                                             uint32_t *echo /*[128][8]*/)
{
    __shared__ __align__(1024) uint8_t btile[32 * 128];      // 32 rows x 64 k x 2 B, 128B swizzle (only k < 16 used)
    __shared__ __align__(1024) uint8_t atile[128 * 128];     // SS variant: A tile 128 rows x 64 k
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 32 * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(btile)[i] = 0;
    for (int i = t; i < 128 * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(atile)[i] = 0;
    __syncthreads();
    if (t < 32)
        for (int k = 0; k < 16; ++k) {
            const int chunk = (k * 2) / 16, within = (k * 2) % 16;
            *reinterpret_cast<uint16_t *>(btile + t * 128 + ((chunk ^ (t & 7)) << 4) + within) = B[t * 16 + k];
        }
    for (int k = 0; k < 16; ++k) {
        const int chunk = (k * 2) / 16, within = (k * 2) % 16;
        *reinterpret_cast<uint16_t *>(atile + t * 128 + ((chunk ^ (t & 7)) << 4) + within) = A[t * 16 + k];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    // A row t -> 8 TMEM columns at column 32
    uint32_t a[8];
    for (int c = 0; c < 8; ++c) a[c] = (uint32_t)A[t * 16 + 2 * c] | ((uint32_t)A[t * 16 + 2 * c + 1] << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tm + lane_sel + 32), "r"(a[0]), "r"(a[1]),
                 "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tm + lane_sel + 32)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) echo[t * 8 + c] = r[c];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (variant > 0) {
        if (t == 0) {
            const uint32_t afmt = variant == 3 ? 1u : 0u, bfmt = (variant == 2 || variant == 3) ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (afmt << 7) | (bfmt << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            auto desc = [](uint32_t addr) {
                uint64_t d = 0;
                d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
                d |= (uint64_t)1 << 16;
                d |= (uint64_t)(1024 >> 4) << 32;
                d |= (uint64_t)1 << 46;
                d |= (uint64_t)2 << 61;
                return d;
            };
            const uint64_t db = desc(smem_u32(btile));
            if (variant == 4) {
                const uint64_t da = desc(smem_u32(atile));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(da),
                             "l"(db), "r"(idesc), "r"(0u)
                             : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm),
                             "r"(tm + 32), "l"(db), "r"(idesc), "r"(0u)
                             : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t d[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
            "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
              "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]),
              "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]),
              "=r"(d[31])
            : "r"(tm + lane_sel)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[t * 32 + j] = __uint_as_float(d[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}

static float h2f(uint16_t h) { __half v; memcpy(&v, &h, 2); return __half2float(v); }
static float b2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    uint16_t hA[128 * 16], hB[32 * 16];
    const bool a_bf = variant == 3, b_bf = variant == 2 || variant == 3;
    srand(1);
    for (int i = 0; i < 128 * 16; ++i) {
        float v = (float)(rand() % 17 - 8) / 4.f;
        if (a_bf) { __nv_bfloat16 x = __float2bfloat16(v); memcpy(&hA[i], &x, 2); } else { __half x = __float2half(v); memcpy(&hA[i], &x, 2); }
    }
    for (int i = 0; i < 32 * 16; ++i) {
        float v = (float)(rand() % 13 - 6) / 2.f;
        if (b_bf) { __nv_bfloat16 x = __float2bfloat16(v); memcpy(&hB[i], &x, 2); } else { __half x = __float2half(v); memcpy(&hB[i], &x, 2); }
    }
    uint16_t *dA, *dB; float *dD; uint32_t *dE;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * 32 * 4); cudaMalloc(&dE, 128 * 8 * 4);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 32 * 4);
    probe<<<1, 128>>>(variant, dA, dB, dD, dE);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
    static float hD[128 * 32]; static uint32_t hE[128 * 8];
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost); cudaMemcpy(hE, dE, sizeof hE, cudaMemcpyDeviceToHost);
    int bad_echo = 0;
    for (int t = 0; t < 128; ++t)
        for (int c = 0; c < 8; ++c)
            if (hE[t * 8 + c] != ((uint32_t)hA[t * 16 + 2 * c] | ((uint32_t)hA[t * 16 + 2 * c + 1] << 16))) ++bad_echo;
    double maxerr = 0;
    if (variant > 0)
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                double acc = 0;
                for (int k = 0; k < 16; ++k) acc += (double)(a_bf ? b2f(hA[m * 16 + k]) : h2f(hA[m * 16 + k])) * (b_bf ? b2f(hB[n * 16 + k]) : h2f(hB[n * 16 + k]));
                maxerr = fmax(maxerr, fabs(acc - hD[m * 32 + n]));
            }
    printf("variant %d: st/ld echo mismatches %d, mma max abs err %.3g  -> %s\n", variant, bad_echo, maxerr, (bad_echo == 0 && maxerr < 1e-3) ? "OK" : "WRONG");
    return 0;
}
