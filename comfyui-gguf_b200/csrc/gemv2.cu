// gemv2.cu -- K3 v2: small-M (M <= 8) Linear on a packed Q4_K / Q5_K weight, bound by reading the packed bytes once.
//
//   Y[m, n] = sum_k X[m, k] * W[n, k] (+ bias),   W[n, k] = D[n, sb] * q[n, k] - Mn[n, sb],   sb = k / 32
//
// Round 1's kernel (gemv.cu, kept as the reference-exact route) materialises every weight in the activation dtype before the
// dot product: ~125 instructions per 16 weights, issue-bound at 0.25 of the HBM peak.  This kernel never forms W:
//   * the 4/5-bit integers go to the tensor core AS INTEGERS: the byte 0x43 (bf16) / 0x64 (fp16) over a quant byte is the exact
//     number 128 + q / 1024 + q, two PRMTs make four of them -- no integer-to-float conversion, no multiply, no cast;
//   * one mma.sync.m16n8k16 sums q-pattern * x over 16 k of ONE sub-block for 16 weight rows x 8 activation rows, fp32
//     accumulate; two of them cover a 32-element sub-block;
//   * the sub-block scale is applied to the 16 x 8 partial sums, not to the weights:
//         acc += D * S - (BIAS * D + Mn) * Xs,      S = sum (BIAS + q) x,   Xs = sum x   over the sub-block (precomputed per CTA)
//     as packed fp32x2 FMAs (FFMA2), with D = fp16(d * sc) and Mn = fp16(dmin * mn) the reference's own sub-block products.
// Per weight that is ~2 instructions (unpack 0.8, scale decode + exchange 0.5, scale application 0.3, MMA + loads 0.4).
// Numerics: the integer unpack and the sub-block products are the reference's; W itself is never rounded, so the result is
// closer to the exact product than the reference's (which rounds W three times in fp16 and once more to bf16) -- this is the
// `fast` contract: within 1e-3 of the reference for fp16 activations, 8e-3 for bf16.
//
// Data movement: a CTA owns 16 consecutive weight rows, which are CONTIGUOUS in the canonical GGUF layout; their packed bytes
// are staged K-chunk by K-chunk (8 super-blocks = 2048 k) into a double-buffered shared-memory tile with one bulk async copy
// (TMA engine, SASS UBLKCP) per row, completion on an mbarrier.  8 warps split a chunk into 16 units of 128 k; their 16 x 8
// partial tiles are reduced through shared memory at the end of a row tile.  Persistent grid.
#include "blocks.cuh"

namespace ggufb200 {

constexpr int kV2Threads = 256;
constexpr int kV2Warps = 8;
constexpr int kV2ChunkBlocks = 8;             // super-blocks per staged chunk (2048 k): two CTAs per SM fit next to the staged activations at K = 3072

template <int TS> struct V2Cfg {
    static constexpr int PITCH = kV2ChunkBlocks * TS + 96;          // row pitch of a staged chunk: == 96 (mod 128), so the four rows of a half-warp's 8-byte loads hit disjoint banks
    static constexpr int BUF = 16 * PITCH;
};

__device__ __forceinline__ unsigned long long v2_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long v2_pack(float lo, float hi)
{
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
template <int ACT> __device__ __forceinline__ void v2_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    if constexpr (ACT == kBF16) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
}
// first MMA of a sub-block: zero accumulator input
template <int ACT> __device__ __forceinline__ void v2_mma0(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    if constexpr (ACT == kBF16) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                     : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                     : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
    }
}
__device__ __forceinline__ uint32_t v2_lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 v2_lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 v2_lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// QK = 4: Q4_K (144-byte super-blocks, qs at +16);  QK = 5: Q5_K (176 bytes, qh at +16, qs at +48)
template <int QK, int ACT, bool XSM>
__global__ void __launch_bounds__(kV2Threads) gemv2_kernel(const uint8_t *__restrict__ W, long long N, long long K, const uint8_t *__restrict__ X,
                                                           long long ldx, int M, const void *__restrict__ bias, int bias_dtype,
                                                           uint8_t *__restrict__ Y, long long ldy)
{
    constexpr int TS = QK == 4 ? 144 : 176;
    constexpr int QS_OFF = QK == 4 ? 16 : 48;
    using Cfg = V2Cfg<TS>;
    // pattern byte and its value: fp16 0x64 -> 1024 + u (and 0x54 -> 64 + q for a high nibble kept in place, Q4_K only); bf16 0x43 -> 128 + u
    constexpr uint32_t MAGIC = ACT == kBF16 ? 0x43434343u : 0x64646464u;
    constexpr float BIAS_LO = ACT == kBF16 ? 128.f : 1024.f;
    constexpr bool INPLACE_HI = (ACT == kF16 && QK == 4);
    constexpr float BIAS_HI = INPLACE_HI ? 64.f : BIAS_LO;

    extern __shared__ __align__(128) uint8_t v2_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(v2_smem);                 // [2]
    float *part = reinterpret_cast<float *>(v2_smem + 64);                   // [8 warps][16][9]
    uint8_t *bufs = v2_smem + 64 + kV2Warps * 16 * 9 * 4 + 64;               // 2 x BUF (16-byte aligned: 64 + 4608 + 64)
    float *xs = reinterpret_cast<float *>(bufs + 2 * Cfg::BUF);              // [K / 32][8]
    // XSM: the activations themselves are staged once per (persistent) CTA: 8 rows, pitch 2K + 64 bytes (rows of a quarter-warp's
    // 16-byte loads then fall into disjoint banks).  Every warp re-reads its K range of X for every row tile: from global memory
    // that is 8 L1 wavefronts per load at M = 8 (8 different rows) and the kernel slows down with M; from shared memory it is 4.
    uint8_t *xsm = reinterpret_cast<uint8_t *>(xs) + (K / 32) * 8 * 4;
    const uint32_t xpitch = (uint32_t)(2 * K + 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const long long row_bytes = K / 256 * TS;
    const int n_blocks = (int)(K / 256);
    const int n_chunks = (n_blocks + kV2ChunkBlocks - 1) / kV2ChunkBlocks;
    const long long n_tiles = (N + 15) / 16;
    const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_steps = my_tiles * n_chunks;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    if constexpr (XSM) {
        const int per_row = (int)(K / 8);                    // 16-byte vectors per activation row
        for (int i = tid; i < 8 * per_row; i += kV2Threads) {
            const int m = i / per_row, v = i - m * per_row;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (m < M) val = *reinterpret_cast<const uint4 *>(X + ((long long)m * ldx + 8ll * v) * 2);
            *reinterpret_cast<uint4 *>(xsm + (size_t)m * xpitch + 16 * v) = val;
        }
    }
    // sub-block sums of the activations: xs[sb][m] = sum of X[m, 32 sb .. 32 sb + 31] (fp32, fixed order); rows >= M read as 0
    for (int i = tid; i < (int)(K / 32) * 8; i += kV2Threads) {
        const int m = i & 7, sb = i >> 3;
        float s = 0.f;
        if (m < M) {
            const uint4 *p = reinterpret_cast<const uint4 *>(X + ((long long)m * ldx + 32ll * sb) * 2);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = p[q];
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if constexpr (ACT == kBF16) s += __uint_as_float(wv[j] << 16) + __uint_as_float(wv[j] & 0xFFFF0000u);
                    else {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&wv[j]));
                        s += f.x + f.y;
                    }
                }
            }
        }
        xs[sb * 8 + m] = s;
    }
    __syncthreads();

    // one elected thread stages (tile, chunk) step `st` into buffer st & 1: one bulk copy per weight row
    auto issue = [&](long long st) {
        const long long tile = blockIdx.x + (st / n_chunks) * gridDim.x;
        const int ch = (int)(st % n_chunks);
        const int cb = min(kV2ChunkBlocks, n_blocks - ch * kV2ChunkBlocks);
        uint8_t *dst = bufs + (st & 1) * Cfg::BUF;
        mbar_arrive_expect_tx(&full[st & 1], (uint32_t)(16 * cb * TS));
#pragma unroll 1
        for (int r = 0; r < 16; ++r) {
            long long n = tile * 16 + r;
            if (n >= N) n = N - 1;                       // a partial last tile re-reads the last row; its results are not stored
            bulk_g2s(dst + r * Cfg::PITCH, W + n * row_bytes + (long long)ch * kV2ChunkBlocks * TS, (uint32_t)(cb * TS), &full[st & 1]);
        }
    };
    if (tid == 0 && n_steps > 0) issue(0);

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // B-fragment column g = activation row g.  A column of B only feeds the same column of D, so lanes whose row does not
    // exist (g >= M) simply read the last valid row: their results are never stored -- no masking in the inner loop.
    const uint8_t *xrow = X + (long long)(g < M ? g : M - 1) * ldx * 2;
    const uint32_t xrow_s = smem_u32(xsm) + (uint32_t)g * xpitch;        // XSM: rows >= M were zero-filled
    const uint32_t xs_base = smem_u32(xs) + (uint32_t)(2 * c) * 4;
    const uint32_t quad_base = (uint32_t)(lane & ~3);

    for (long long st = 0; st < n_steps; ++st) {
        const int ch = (int)(st % n_chunks);
        const int cb = min(kV2ChunkBlocks, n_blocks - ch * kV2ChunkBlocks);
        if (tid == 0 && st + 1 < n_steps) issue(st + 1);          // the other buffer was released by the __syncthreads of step st - 1
        mbar_wait(&full[st & 1], (uint32_t)((st >> 1) & 1));
        const uint32_t buf = smem_u32(bufs + (st & 1) * Cfg::BUF);
        const uint32_t rg = buf + (uint32_t)(g * Cfg::PITCH), rg8 = buf + (uint32_t)((g + 8) * Cfg::PITCH);

        for (int u = warp; u < 2 * cb; u += kV2Warps) {
            const int bl = u >> 1, h = u & 1;                       // super-block inside the chunk, 128-element half
            const uint32_t bg = rg + (uint32_t)(bl * TS), bg8 = rg8 + (uint32_t)(bl * TS);
            const long long kblk = ((long long)ch * kV2ChunkBlocks + bl) * 256;
            // ---- sub-block scales: thread c of a quad decodes sub-block j = 4h + c for both rows, the quad exchanges by shuffle
            float Dg, Eg, Dg8, Eg8;
            {
                const uint4 hg = v2_lds128(bg), hg8 = v2_lds128(bg8);
                const int j = 4 * h + c;
                const int sh = 8 * (j & 3);
                auto dec = [&](const uint4 &hd, float &D, float &E) {
                    const uint32_t a = (hd.y >> sh) & 0xFFu, b = (hd.z >> sh) & 0xFFu, cc = (hd.w >> sh) & 0xFFu;
                    const uint32_t sc = h ? ((cc & 0x0Fu) | ((a >> 6) << 4)) : (a & 63u);
                    const uint32_t mn = h ? ((cc >> 4) | ((b >> 6) << 4)) : (b & 63u);
                    uint32_t scm = sc | (mn << 16) | 0x64006400u;
                    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400u));
                    const __half2 v = __hsub2_rn(*reinterpret_cast<__half2 *>(&scm), k1024);
                    uint32_t dm = hd.x;
                    const float2 DM = __half22float2(__hmul2_rn(*reinterpret_cast<__half2 *>(&dm), v));     // fp16(d*sc), fp16(dmin*mn): the reference's products
                    D = DM.x;
                    E = -fmaf((c & 1) ? BIAS_HI : BIAS_LO, DM.x, DM.y);         // stored negated: the update is acc += D * S + E * Xs
                };
                dec(hg, Dg, Eg);
                dec(hg8, Dg8, Eg8);
            }
            uint32_t qhg[2] = {0, 0}, qhg8[2] = {0, 0};
            if constexpr (QK == 5) {
                const uint2 a = v2_lds64(bg + 16 + 8 * c), b = v2_lds64(bg8 + 16 + 8 * c);
                qhg[0] = a.x; qhg[1] = a.y; qhg8[0] = b.x; qhg8[1] = b.y;
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const uint2 wg = v2_lds64(bg + QS_OFF + 32 * (2 * h + p) + 8 * c), wg8 = v2_lds64(bg8 + QS_OFF + 32 * (2 * h + p) + 8 * c);
#pragma unroll
                for (int odd = 0; odd < 2; ++odd) {
                    const int s = 2 * p + odd;                      // sub-block 4h + s
                    const int sb = (int)(kblk >> 5) + 4 * h + s;
                    const uint32_t src = quad_base | (uint32_t)s;
                    const float D0 = __shfl_sync(0xffffffffu, Dg, src), E0 = __shfl_sync(0xffffffffu, Eg, src);
                    const float D8 = __shfl_sync(0xffffffffu, Dg8, src), E8 = __shfl_sync(0xffffffffu, Eg8, src);
                    uint4 xv;
                    if constexpr (XSM) xv = v2_lds128(xrow_s + (uint32_t)((kblk + 32 * (4 * h + s) + 8 * c) * 2));
                    else xv = *reinterpret_cast<const uint4 *>(xrow + (kblk + 32 * (4 * h + s) + 8 * c) * 2);
                    float d[4];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        uint32_t ug = i ? wg.y : wg.x, ug8 = i ? wg8.y : wg8.x;
                        uint32_t magic = MAGIC;
                        if (odd) {
                            if constexpr (INPLACE_HI) {
                                ug &= 0xF0F0F0F0u; ug8 &= 0xF0F0F0F0u;
                                magic = 0x54545454u;
                            } else {
                                ug = (ug >> 4) & 0x0F0F0F0Fu; ug8 = (ug8 >> 4) & 0x0F0F0F0Fu;
                            }
                        } else {
                            ug &= 0x0F0F0F0Fu; ug8 &= 0x0F0F0F0Fu;
                        }
                        if constexpr (QK == 5) {
                            const int jj = 4 * h + s;
                            ug |= ((qhg[i] >> jj) & 0x01010101u) << 4;
                            ug8 |= ((qhg8[i] >> jj) & 0x01010101u) << 4;
                        }
                        if (i == 0)
                            v2_mma0<ACT>(d, prmt(ug, magic, 0x4140u), prmt(ug8, magic, 0x4140u), prmt(ug, magic, 0x4342u), prmt(ug8, magic, 0x4342u), xv.x, xv.y);
                        else
                            v2_mma<ACT>(d, prmt(ug, magic, 0x4140u), prmt(ug8, magic, 0x4140u), prmt(ug, magic, 0x4342u), prmt(ug8, magic, 0x4342u), xv.z, xv.w);
                    }
                    // acc += D * S - E * Xs   (packed fp32 pairs: accumulators (0,1) belong to row g, (2,3) to row g + 8; tokens 2c, 2c + 1)
                    const uint2 xsv = v2_lds64(xs_base + (uint32_t)sb * 32);
                    const unsigned long long xs2 = (unsigned long long)xsv.x | ((unsigned long long)xsv.y << 32);
                    unsigned long long a01 = v2_pack(acc[0], acc[1]), a23 = v2_pack(acc[2], acc[3]);
                    a01 = v2_fma2(v2_pack(D0, D0), v2_pack(d[0], d[1]), a01);
                    a01 = v2_fma2(v2_pack(E0, E0), xs2, a01);
                    a23 = v2_fma2(v2_pack(D8, D8), v2_pack(d[2], d[3]), a23);
                    a23 = v2_fma2(v2_pack(E8, E8), xs2, a23);
                    acc[0] = __uint_as_float((uint32_t)a01); acc[1] = __uint_as_float((uint32_t)(a01 >> 32));
                    acc[2] = __uint_as_float((uint32_t)a23); acc[3] = __uint_as_float((uint32_t)(a23 >> 32));
                }
            }
        }

        if (ch == n_chunks - 1) {
            // ---- row tile complete: reduce the 8 warps' 16 x 8 partials, add the bias, store
            float(*pt)[16][9] = reinterpret_cast<float(*)[16][9]>(part);
            pt[warp][g][2 * c] = acc[0];
            pt[warp][g][2 * c + 1] = acc[1];
            pt[warp][g + 8][2 * c] = acc[2];
            pt[warp][g + 8][2 * c + 1] = acc[3];
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
            __syncthreads();
            if (tid < 128) {
                const int f = tid >> 3, m = tid & 7;
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < kV2Warps; ++w) a += pt[w][f][m];
                const long long n = (blockIdx.x + (st / n_chunks) * gridDim.x) * 16 + f;
                if (n < N && m < M) {
                    if (bias) {
                        float b;
                        if (bias_dtype == kF32) b = reinterpret_cast<const float *>(bias)[n];
                        else if (bias_dtype == kF16) b = __half2float(reinterpret_cast<const __half *>(bias)[n]);
                        else b = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(bias)[n]);
                        if constexpr (ACT == kBF16) a += __bfloat162float(__float2bfloat16_rn(b));      // ops.py:205-207: bias is cast to x.dtype first
                        else a += __half2float(__float2half_rn(b));
                    }
                    if constexpr (ACT == kBF16) reinterpret_cast<__nv_bfloat16 *>(Y)[(long long)m * ldy + n] = __float2bfloat16_rn(a);
                    else reinterpret_cast<__half *>(Y)[(long long)m * ldy + n] = __float2half_rn(a);
                }
            }
        }
        __syncthreads();          // every warp is done with buffer st & 1 (and with `part`)
    }
}

template <int QK, int ACT, bool XSM>
static int gemv2_launch2(const void *W, long long N, long long K, const void *X, long long M, long long ldx, const void *bias, int bias_dtype, void *Y,
                         long long ldy, int smem, cudaStream_t st)
{
    auto kern = gemv2_kernel<QK, ACT, XSM>;
    static unsigned char attr[64] = {};
    if (!ensure_dynamic_smem(kern, 227 * 1024, attr)) return GGUFB200_E_CUDA;      // the size depends on K: raise the cap once, to the maximum
    const long long tiles = (N + 15) / 16;
    int per_sm = (227 * 1024) / (smem + 1024);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)sm_count() * per_sm;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    kern<<<grid, kV2Threads, smem, st>>>(reinterpret_cast<const uint8_t *>(W), N, K, reinterpret_cast<const uint8_t *>(X), ldx, (int)M, bias, bias_dtype,
                                         reinterpret_cast<uint8_t *>(Y), ldy);
    return cudaGetLastError() == cudaSuccess ? GGUFB200_OK : GGUFB200_E_CUDA;
}

template <int QK, int ACT>
static int gemv2_launch(const void *W, long long N, long long K, const void *X, long long M, long long ldx, const void *bias, int bias_dtype, void *Y,
                        long long ldy, cudaStream_t st)
{
    constexpr int TS = QK == 4 ? 144 : 176;
    const int base = 64 + kV2Warps * 16 * 9 * 4 + 64 + 2 * V2Cfg<TS>::BUF + (int)(K / 32) * 8 * 4;
    const long long with_x = base + 8 * (2 * K + 64);
    if (base > 227 * 1024) return GGUFB200_E_UNSUPPORTED;
    // stage X when two CTAs per SM still fit (K <= ~3.4 k for Q4_K): beyond that the latency hiding of the second CTA is worth more
    if (2 * (with_x + 1024) <= 227 * 1024) return gemv2_launch2<QK, ACT, true>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, (int)with_x, st);
    return gemv2_launch2<QK, ACT, false>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, base, st);
}

bool gemv2_supported(int type, const void *W, long long N, long long K, long long M)
{
    if (type != T_Q4_K && type != T_Q5_K) return false;
    if (M < 1 || M > 8 || K % 256 != 0 || N < 1) return false;
    return (reinterpret_cast<uintptr_t>(W) & 15) == 0;          // rows of whole 144 / 176-byte super-blocks are then 16-byte aligned
}

int gemv2_dispatch(int type, const void *W, long long N, long long K, const void *X, long long M, long long ldx, int act_dtype, const void *bias,
                   int bias_dtype, void *Y, long long ldy, cudaStream_t st)
{
    if (!gemv2_supported(type, W, N, K, M)) return GGUFB200_E_UNSUPPORTED;
    if (type == T_Q4_K)
        return act_dtype == kBF16 ? gemv2_launch<4, kBF16>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, st)
                                  : gemv2_launch<4, kF16>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, st);
    return act_dtype == kBF16 ? gemv2_launch<5, kBF16>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, st)
                              : gemv2_launch<5, kF16>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, st);
}

}  // namespace ggufb200
