// probe_write.cu -- micro-benchmark: how fast can a B200 WRITE to HBM, and with which store flavour?
// (roofline sanity for the write-dominated dequant kernel; not part of the library)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE> __global__ void __launch_bounds__(256) fill_kernel(uint4 *dst, size_t n16, uint64_t policy)
{
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) {
        if (MODE == 0) dst[i] = v;
        else if (MODE == 1) asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        else if (MODE == 2) asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
        else if (MODE == 3) asm volatile("st.global.wt.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        else if (MODE == 4) asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

// tile-contiguous variant: each CTA writes whole 16 KiB tiles (like the dequant kernel)
__global__ void __launch_bounds__(256) fill_tiles(uint4 *dst, size_t n16)
{
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    const size_t tiles = n16 / 1024;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x)
#pragma unroll
        for (int p = 0; p < 4; ++p) dst[t * 1024 + p * 256 + threadIdx.x] = v;
}

// TMA bulk store: smem -> global, 16 KiB per CTA iteration
__global__ void __launch_bounds__(256) fill_tma(uint8_t *dst, size_t bytes, int chunk)
{
    extern __shared__ __align__(128) uint8_t buf[];
    for (int i = threadIdx.x; i < chunk / 16; i += 256) reinterpret_cast<uint4 *>(buf)[i] = make_uint4(i, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t n = bytes / chunk;
        for (size_t t = blockIdx.x; t < n; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * chunk), "r"(smem_u32(buf)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main()
{
    const size_t sizes[2] = {132120576ull, 1056964608ull};
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (size_t bytes : sizes) {
        const int NB = 3;
        std::vector<uint8_t *> bufs(NB);
        for (auto &p : bufs) cudaMalloc(&p, bytes);
        uint64_t pol_first = 0, pol_last = 0;
        {
            uint64_t *d; cudaMalloc(&d, 16);
            // policies are created on the device in a tiny kernel-free way: use createpolicy in a lambda kernel
            (void)d;
        }
        auto run = [&](const char *name, auto launch) {
            for (int w = 0; w < 3; ++w) launch(bufs[w % NB]);
            cudaDeviceSynchronize();
            cudaEventRecord(a);
            const int iters = 12;
            for (int i = 0; i < iters; ++i) launch(bufs[i % NB]);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            printf("%10zu B  %-34s %8.1f us  %8.1f GB/s %s\n", bytes, name, ms / iters * 1e3, bytes / (ms / iters * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        const size_t n16 = bytes / 16;
        for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
            char nm[64];
            snprintf(nm, 64, "st.v4 default grid=%d", g); run(nm, [&](uint8_t *p) { fill_kernel<0><<<g, 256>>>((uint4 *)p, n16, 0); });
        }
        run("st.v4 default grid=n16/256", [&](uint8_t *p) { fill_kernel<0><<<(unsigned)(n16 / 256), 256>>>((uint4 *)p, n16, 0); });
        run("st.cs.v4 grid=1184", [&](uint8_t *p) { fill_kernel<1><<<1184, 256>>>((uint4 *)p, n16, 0); });
        run("st.wt.v4 grid=1184", [&](uint8_t *p) { fill_kernel<3><<<1184, 256>>>((uint4 *)p, n16, 0); });
        run("st.L1::no_allocate grid=1184", [&](uint8_t *p) { fill_kernel<4><<<1184, 256>>>((uint4 *)p, n16, 0); });
        run("tile-contiguous 16KiB grid=888", [&](uint8_t *p) { fill_tiles<<<888, 256>>>((uint4 *)p, n16); });
        run("tile-contiguous 16KiB grid=592", [&](uint8_t *p) { fill_tiles<<<592, 256>>>((uint4 *)p, n16); });
        for (int chunk : {4096, 16384, 65536}) {
            char nm[64];
            snprintf(nm, 64, "TMA bulk store chunk=%d grid=592", chunk);
            cudaFuncSetAttribute(fill_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            run(nm, [&](uint8_t *p) { fill_tma<<<592, 256, chunk>>>(p, bytes, chunk); });
        }
        run("cudaMemsetAsync", [&](uint8_t *p) { cudaMemsetAsync(p, 1, bytes); });
        run("cudaMemcpyAsync D2D (r+w bytes/2)", [&](uint8_t *p) { cudaMemcpyAsync(p, bufs[(p == bufs[0]) ? 1 : 0], bytes, cudaMemcpyDeviceToDevice); });
        for (auto p : bufs) cudaFree(p);
        (void)pol_first; (void)pol_last;
    }
    return 0;
}
