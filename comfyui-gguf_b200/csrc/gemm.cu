// gemm.cu -- K2: large-M Linear on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   Y[M,N] = X[M,K] * W[N,K]^T (+ bias)      X, Y in the activation dtype (fp16 / bf16), fp32 accumulate
//
// One warp-specialised kernel, two B-operand producers:
//   * FUSED  (Q = a Block<> type): W stays PACKED in HBM.  Eight dequant warps read the packed rows of the
//     CTA's N-tile straight from global/L2, unpack them with blocks.cuh (the reference's rounding sequence,
//     then the cast to the activation dtype -- the exact operand the reference hands to F.linear), and write
//     the K-major 128B-swizzled B tile into shared memory; fp16/bf16 W is never materialised in HBM.
//     Replaces ops.py:242-244 (dequant chain + F.linear).
//   * DENSE  (Q = void): B tiles come from a 2-D TMA tensor map over an already dense W[N,K]
//     (F16/BF16 Linears, and the second half of GGUFB200_ALGO_DEQUANT_MMA).
//
// CTA tile 256 (M) x BN (N) x 64 (K): two UMMA 128xBNx16 accumulators share every B tile, so each dequantised
// element feeds 512 flops.  512 threads:
//   warp 0        TMA producer (A tile 256x64 via cp.async.bulk.tensor, 128B swizzle; B too in DENSE mode)
//   warp 1        MMA issuer (one elected lane: 8 tcgen05.mma per k-block, tcgen05.commit -> empty barrier)
//   warp 2        TMEM allocate / deallocate
//   warps 4-7     epilogue for accumulator 0  (tcgen05.ld -> +bias -> cast -> st.global)
//   warps 8-15    FUSED: dequant producers during the main loop; warps 8-11 then run the epilogue of accumulator 1
// Pipelines: smem ring full_a/full_b/empty (mbarrier), TMEM full (tcgen05.commit).
#include <cuda.h>

#include "blocks.cuh"

namespace ggufb200 {

constexpr int kGemmThreads = 512;
constexpr int kBM = 256;
constexpr int kBK = 64;
constexpr int kDequantThreads = 256;

template <int BN> struct GemmCfg {
    static constexpr int STAGES = BN == 256 ? 3 : 4;
    static constexpr int A_BYTES = kBM * kBK * 2;
    static constexpr int B_BYTES = BN * kBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KiB alignment of the tiles
    static constexpr int TMEM_COLS = 2 * BN;                              // 256 or 512 (power of two)
};

// ------------------------------------------------------------------ tcgen05 / TMA PTX
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset = 1024 B between 8-row atoms, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::f16: D=f32, A/B = fp16 (0) or bf16 (1), both K-major, M=128, N=BN
template <int ACT, int BN> __device__ __forceinline__ constexpr uint32_t umma_idesc()
{
    uint32_t fmt = ACT == kBF16 ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int ACT> __device__ __forceinline__ float bias_value(const void *bias, int bias_dtype, long long n)
{
    float b;
    if (bias_dtype == kF32) b = reinterpret_cast<const float *>(bias)[n];
    else if (bias_dtype == kF16) b = __half2float(reinterpret_cast<const __half *>(bias)[n]);
    else b = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(bias)[n]);
    // the reference casts the bias to the activation dtype before F.linear (ops.py:205-207)
    if constexpr (ACT == kBF16) return __bfloat162float(__float2bfloat16_rn(b));
    else return __half2float(__float2half_rn(b));
}

template <int ACT> __device__ __forceinline__ uint32_t pack_act(float a, float b)
{
    if constexpr (ACT == kBF16) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&v);
    } else {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&v);
    }
}

struct GemmParams {
    const uint8_t *W;      // FUSED: packed rows
    long long row_bytes;   // FUSED: bytes per packed row
    long long M, N, K;
    const void *bias;
    int bias_dtype;
    uint8_t *Y;
    long long ldy;
    int tiles_m;
};

template <class Q, int MATH, int ACT, int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p)
{
    using Cfg = GemmCfg<BN>;
    constexpr bool FUSED = !std::is_same<Q, void>::value;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + STAGES * Cfg::STAGE_BYTES);
    uint64_t *full_a = bars;                 // [STAGES] TMA landed (A, and B in DENSE mode)
    uint64_t *full_b = bars + STAGES;        // [STAGES] dequant warps finished the B tile (FUSED)
    uint64_t *empty = bars + 2 * STAGES;     // [STAGES] MMAs that read the slot have completed
    uint64_t *tmem_full = bars + 3 * STAGES; // accumulators complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x % p.tiles_m;
    const int tile_n = blockIdx.x / p.tiles_m;
    const long long m0 = (long long)tile_m * kBM;
    const long long n0 = (long long)tile_n * BN;
    const int num_kb = (int)(p.K / kBK);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_a[s], 1);
            mbar_init(&full_b[s], kDequantThreads);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty[s], (uint32_t)(((kb / STAGES) & 1) ^ 1));
                uint8_t *a_dst = tiles + s * Cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(&full_a[s], FUSED ? Cfg::A_BYTES : Cfg::STAGE_BYTES);
                tma_load_2d(a_dst, &tmA, &full_a[s], kb * kBK, (int)m0);
                if constexpr (!FUSED) tma_load_2d(a_dst + Cfg::A_BYTES, &tmB, &full_a[s], kb * kBK, (int)n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc<ACT, BN>();
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t par = (uint32_t)((kb / STAGES) & 1);
                mbar_wait(&full_a[s], par);
                if constexpr (FUSED) mbar_wait(&full_b[s], par);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(tiles + s * Cfg::STAGE_BYTES);
                const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                for (int j = 0; j < kBK / 16; ++j) {
                    const uint64_t db = umma_desc_sw128(b_addr + j * 32);
                    const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
                    umma_f16(tmem_base, umma_desc_sw128(a_addr + j * 32), db, idesc, acc);
                    umma_f16(tmem_base + BN, umma_desc_sw128(a_addr + 128 * 128 + j * 32), db, idesc, acc);
                }
                umma_commit(&empty[s]);   // implicit tcgen05.fence::before_thread_sync
            }
            umma_commit(tmem_full);
        }
    } else if (warp >= 8) {
        // ===================== dequant producers (FUSED only)
        if constexpr (FUSED) {
            const int t = threadIdx.x - 256;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty[s], (uint32_t)(((kb / STAGES) & 1) ^ 1));
                const uint32_t b_base = smem_u32(tiles + s * Cfg::STAGE_BYTES + Cfg::A_BYTES);
#pragma unroll 2
                for (int i = 0; i < BN * 8 / kDequantThreads; ++i) {
                    const int task = i * kDequantThreads + t;
                    const int row = task >> 3, chunk = task & 7;
                    const uint32_t dst = b_base + row * 128 + ((chunk ^ (row & 7)) << 4);
                    if (n0 + row < p.N) {
                        const long long k = (long long)kb * kBK + chunk * 8;
                        const uint8_t *blk = p.W + (n0 + row) * p.row_bytes + (k / Q::BS) * Q::TS;
                        typename Math<MATH>::T2 v[4];
                        dequant_run<Q, MATH, 8>(blk, (int)(k % Q::BS), v);
                        st_shared_v4(dst, pack16<ACT, MATH>(v[0]), pack16<ACT, MATH>(v[1]), pack16<ACT, MATH>(v[2]), pack16<ACT, MATH>(v[3]));
                    } else {
                        st_shared_v4(dst, 0, 0, 0, 0);
                    }
                }
                fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
                mbar_arrive(&full_b[s]);
            }
        }
    }

    // ===================== epilogue: warps 4-7 -> accumulator 0, warps 8-11 -> accumulator 1
    if (warp >= 4 && warp < 12) {
        const int acc = (warp >= 8) ? 1 : 0;
        const int quad = warp & 3;   // TMEM lane quadrant this warp may access
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const long long m = m0 + acc * 128 + quad * 32 + lane;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
        uint8_t *yrow = p.Y + (m * p.ldy + n0) * 2;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr0 + c0, r);
            tmem_ld_wait();
            if (m < p.M) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const long long n = n0 + c0 + g * 8;
                    if (n < p.N) {   // N % 8 == 0 is validated on the host
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v0 = __uint_as_float(r[g * 8 + 2 * j]), v1 = __uint_as_float(r[g * 8 + 2 * j + 1]);
                            if (p.bias) {
                                v0 += bias_value<ACT>(p.bias, p.bias_dtype, n + 2 * j);
                                v1 += bias_value<ACT>(p.bias, p.bias_dtype, n + 2 * j + 1);
                            }
                            o[j] = pack_act<ACT>(v0, v1);
                        }
                        st_global_v4(yrow + (c0 + g * 8) * 2, o[0], o[1], o[2], o[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D row-major [rows, K] tensor of 16-bit elements, box = 64 (K) x box_rows, 128-byte swizzle, zero fill out of bounds
static bool make_map(CUtensorMap *tm, const void *base, long long rows, long long K, long long ld, int act, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = act == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    return fn(tm, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_bn(long long M, long long N)
{
    const int sms = sm_count();
    long long tm = (M + kBM - 1) / kBM;
    auto eff = [&](int bn) {
        long long tiles = tm * ((N + bn - 1) / bn);
        long long waves = (tiles + sms - 1) / sms;
        return (double)tiles / (double)(waves * sms);
    };
    // 256-wide tiles halve the X traffic per flop; take them unless the last wave would be much emptier
    return eff(256) + 0.08 >= eff(128) ? 256 : 128;
}

template <class Q, int MATH, int ACT, int BN>
static int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, cudaStream_t st)
{
    using Cfg = GemmCfg<BN>;
    auto kern = gemm_kernel<Q, MATH, ACT, BN>;
    static unsigned char attr[64] = {};
    if (!ensure_dynamic_smem(kern, Cfg::SMEM, attr)) return GGUFB200_E_CUDA;
    GemmParams q = p;
    q.tiles_m = (int)((p.M + kBM - 1) / kBM);
    long long tiles_n = (p.N + BN - 1) / BN;
    kern<<<(unsigned)(q.tiles_m * tiles_n), kGemmThreads, Cfg::SMEM, st>>>(tmA, tmB, q);
    return cudaGetLastError() == cudaSuccess ? GGUFB200_OK : GGUFB200_E_CUDA;
}

template <class Q, int ACT>
static int gemm_fused_act(const void *W, long long N, long long K, const void *X, long long M, long long ldx, const void *bias, int bias_dtype,
                          void *Y, long long ldy, cudaStream_t st)
{
    CUtensorMap tmA;
    if (!make_map(&tmA, X, M, K, ldx, ACT, kBM)) return GGUFB200_E_CUDA;
    GemmParams p{};
    p.W = reinterpret_cast<const uint8_t *>(W);
    p.row_bytes = K / Q::BS * Q::TS;
    p.M = M; p.N = N; p.K = K;
    p.bias = bias; p.bias_dtype = bias_dtype;
    p.Y = reinterpret_cast<uint8_t *>(Y); p.ldy = ldy;
    if (pick_bn(M, N) == 256) return launch_gemm<Q, kF16, ACT, 256>(tmA, tmA, p, st);
    return launch_gemm<Q, kF16, ACT, 128>(tmA, tmA, p, st);
}

int gemm_fused_supported(int type)
{
    switch (type) {
    case T_Q4_0: case T_Q4_1: case T_Q5_0: case T_Q5_1: case T_Q8_0: case T_Q2_K: case T_Q3_K: case T_Q4_K: case T_Q5_K:
    case T_Q6_K: case T_IQ4_NL: case T_IQ4_XS:
        return 1;
    }
    return 0;
}

// fp16 reference math only (the default `dequant_dtype=None`); other math dtypes take GGUFB200_ALGO_DEQUANT_MMA
int gemm_fused_dispatch(int type, const void *W, long long N, long long K, const void *X, long long M, long long ldx, int act_dtype,
                        int math_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st)
{
    if (math_dtype != kF16 || K % kBK != 0 || N % 8 != 0) return GGUFB200_E_UNSUPPORTED;
#define GGUFB200_FUSED_CASE(T)                                                                                              \
    case T:                                                                                                                 \
        return act_dtype == kBF16 ? gemm_fused_act<Block<T>, kBF16>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, st)        \
                                  : gemm_fused_act<Block<T>, kF16>(W, N, K, X, M, ldx, bias, bias_dtype, Y, ldy, st);
    switch (type) {
        GGUFB200_FUSED_CASE(T_Q4_0)
        GGUFB200_FUSED_CASE(T_Q4_1)
        GGUFB200_FUSED_CASE(T_Q5_0)
        GGUFB200_FUSED_CASE(T_Q5_1)
        GGUFB200_FUSED_CASE(T_Q8_0)
        GGUFB200_FUSED_CASE(T_Q2_K)
        GGUFB200_FUSED_CASE(T_Q3_K)
        GGUFB200_FUSED_CASE(T_Q4_K)
        GGUFB200_FUSED_CASE(T_Q5_K)
        GGUFB200_FUSED_CASE(T_Q6_K)
        GGUFB200_FUSED_CASE(T_IQ4_NL)
        GGUFB200_FUSED_CASE(T_IQ4_XS)
    }
#undef GGUFB200_FUSED_CASE
    return GGUFB200_E_UNSUPPORTED;
}

int gemm_dense_dispatch(const void *W, long long N, long long K, long long ldw, const void *X, long long M, long long ldx, int act_dtype,
                        const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st)
{
    if (K % kBK != 0 || N % 8 != 0) return GGUFB200_E_UNSUPPORTED;
    int bn = pick_bn(M, N);
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, X, M, K, ldx, act_dtype, kBM)) return GGUFB200_E_CUDA;
    if (!make_map(&tmB, W, N, K, ldw, act_dtype, bn)) return GGUFB200_E_CUDA;
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = bias; p.bias_dtype = bias_dtype;
    p.Y = reinterpret_cast<uint8_t *>(Y); p.ldy = ldy;
    if (act_dtype == kBF16) {
        if (bn == 256) return launch_gemm<void, kF16, kBF16, 256>(tmA, tmB, p, st);
        return launch_gemm<void, kF16, kBF16, 128>(tmA, tmB, p, st);
    }
    if (bn == 256) return launch_gemm<void, kF16, kF16, 256>(tmA, tmB, p, st);
    return launch_gemm<void, kF16, kF16, 128>(tmA, tmB, p, st);
}

}  // namespace ggufb200
