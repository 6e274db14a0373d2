// dequant.cu -- K1: standalone GGUF block dequant (HBM-bound streaming kernel).
//
// Replaces dequant.py:30-44 + every dequantize_blocks_* (dequant.py:61-285) + the final
// `.to(dtype)` (dequant.py:23) with ONE kernel launch per tensor.
//
// Data movement (algorithmic bytes per element = TS/BS read + sizeof(out) written):
//   * the packed block stream is treated as a flat byte stream (block sizes 18/22/34/84/110/
//     210 B are not 16 B multiples, so 2-D tensor maps are illegal for most shapes); it is cut
//     into tiles of TILE_ELEMS elements whose byte span is always a multiple of 16 B
//   * each tile is staged into shared memory by ONE elected thread with the TMA engine
//     (cp.async.bulk, SASS UBLKCP) into a STAGES-deep ring, completion on an mbarrier;
//     the ring is refilled as soon as a slot has been consumed
//   * every thread unpacks runs of 4/8 consecutive elements from shared memory (blocks.cuh)
//     and writes one 16-byte vector per run, so a warp stores 512 contiguous bytes
//   * persistent grid: gridDim = min(tiles, SMs * CTAs/SM), tile = blockIdx + i * gridDim
#include "blocks.cuh"

namespace ggufb200 {

constexpr int kThreads = 256;

int g_dequant_ctas_per_sm = 0;  // 0 = default; set through ggufb200_set_tuning(0, v)
int g_dequant_pdl = 1;          // programmatic dependent launch of the dequant kernel; ggufb200_set_tuning(1, 0/1)

// bulk async copy shared -> global (TMA engine), tracked with bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_le1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One thread = one run of 32 consecutive elements (a whole 32-block, or one/two scale groups of a K-quant
// super-block): the header is decoded once per 32 elements.  The 32 results (64 B fp16/bf16, 128 B fp32) go to a
// linear output tile in shared memory in 16-byte chunks -- chunk order rotated per lane so the STS.128 are bank
// conflict free -- and the finished tile leaves through ONE bulk async store (TMA engine), so the threads never
// compute a global address and every HBM write is a full line.
template <class Q, int MATH, int OUT, int STAGES, int TILE_ELEMS>
__global__ void __launch_bounds__(kThreads) dequant_kernel(const uint8_t *__restrict__ src, void *__restrict__ dst, long long n_blocks,
                                                           int bulk_ok)
{
    constexpr int OB = OutT<OUT>::bytes;
    constexpr int EPC = 16 / OB;                           // elements per 16-byte chunk: 8 or 4
    constexpr int CH = 32 / EPC;                           // chunks per thread: 4 or 8
    constexpr int TILE_BLOCKS = TILE_ELEMS / Q::BS;
    constexpr int TILE_BYTES = TILE_BLOCKS * Q::TS;
    constexpr int SLOT_BYTES = TILE_BYTES + 16;            // +16: tail over-read rounding
    constexpr int OUT_BYTES = TILE_ELEMS * OB;
    constexpr int GROUP = GroupOf<Q>::value;
    static_assert(TILE_BYTES % 16 == 0, "tile byte span must be a multiple of 16");
    static_assert(TILE_ELEMS == kThreads * 32 && TILE_ELEMS % Q::BS == 0, "tile shape");

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);   // STAGES mbarriers
    uint8_t *slots = smem + 128;
    uint8_t *outs = slots + STAGES * SLOT_BYTES;           // two output tiles

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const long long n_tiles = (n_blocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    const long long total_bytes = n_blocks * (long long)Q::TS;
    const long long n_elems = n_blocks * (long long)Q::BS;

    // Programmatic dependent launch: let the next kernel in the stream be scheduled while this one drains, and
    // do this kernel's own set-up before waiting for the previous kernel's memory to be complete and visible.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // bytes of tile t (the last tile may be short); bulk copies are rounded up to 16 B, which
    // stays inside the 16-byte granule that holds the last valid byte
    auto issue = [&](long long t, int slot) {
        long long off = t * (long long)TILE_BYTES;
        long long len = total_bytes - off;
        if (len > TILE_BYTES) len = TILE_BYTES;
        uint32_t bytes = (uint32_t)((len + 15) & ~15LL);
        mbar_arrive_expect_tx(&full[slot], bytes);
        bulk_g2s(slots + slot * SLOT_BYTES, src + off, bytes, &full[slot]);
    };

    if (bulk_ok && tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_tiles) issue(t, s);
        }
    }

    // this thread's run inside any tile
    const int blk_in_tile = (tid * 32) / Q::BS;
    const int e0 = (tid * 32) % Q::BS;
    const int rot = (CH == 4) ? (lane >> 1) : lane;

    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int slot = it % STAGES;
        const uint8_t *tile = slots + slot * SLOT_BYTES;
        uint8_t *otile = outs + (it & 1) * OUT_BYTES;
        if (bulk_ok) {
            mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
        } else {
            // unaligned source pointer (never produced by torch allocations): cooperative byte copy
            long long off = t * (long long)TILE_BYTES;
            long long len = total_bytes - off;
            if (len > TILE_BYTES) len = TILE_BYTES;
            for (int i = tid; i < (int)len; i += kThreads) slots[slot * SLOT_BYTES + i] = src[off + i];
            __syncthreads();
        }

        const long long elem_base = t * (long long)TILE_ELEMS;
        const long long left = n_elems - elem_base;
        const int tile_elems = left < TILE_ELEMS ? (int)left : TILE_ELEMS;   // a multiple of 32
        if (tid * 32 < tile_elems) {
            const uint8_t *blk = tile + blk_in_tile * Q::TS;
            const GroupScale<MATH> g0 = group_scale<Q, MATH>(blk, e0);
            GroupScale<MATH> g1 = g0;
            if constexpr (GROUP == 16) g1 = group_scale<Q, MATH>(blk, e0 + 16);
            const uint32_t obase = smem_u32(otile) + tid * (32 * OB);
#pragma unroll
            for (int p = 0; p < CH; ++p) {
                const int c = (p + rot) & (CH - 1);          // rotated chunk order: conflict-free STS.128
                const int e = e0 + c * EPC;
                const bool second = (GROUP == 16) && (c * EPC >= 16);
                GroupScale<MATH> g;
                g.a = second ? g1.a : g0.a;
                g.b = second ? g1.b : g0.b;
                typename Math<MATH>::T2 v[EPC / 2];
                dequant_elems<Q, MATH, EPC>(blk, e, g, v);
                if constexpr (OUT == kF32) {
                    float2 f0 = Math<MATH>::to_f32x2(v[0]), f1 = Math<MATH>::to_f32x2(v[1]);
                    st_shared_v4(obase + c * 16, __float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f1.x), __float_as_uint(f1.y));
                } else {
                    st_shared_v4(obase + c * 16, pack16<OUT, MATH>(v[0]), pack16<OUT, MATH>(v[1]), pack16<OUT, MATH>(v[2]),
                                 pack16<OUT, MATH>(v[3]));
                }
            }
        }
        fence_proxy_async_smem();   // output tile written through the generic proxy, read by the TMA engine
        __syncthreads();            // tile complete; every thread is also done reading the input slot
        if (tid == 0) {
            bulk_s2g(reinterpret_cast<uint8_t *>(dst) + elem_base * OB, otile, (uint32_t)(tile_elems * OB));
            if (bulk_ok) {
                long long tn = t + (long long)STAGES * gridDim.x;
                if (tn < n_tiles) issue(tn, slot);
            }
            bulk_wait_read_le1();   // the other output buffer (stored one tile ago) has been read out
        }
        __syncthreads();
    }
    if (tid == 0) bulk_wait_all();
}

// BF16 "quantised" type (dequant.py:61-62): widen to fp32, then cast to the output dtype
template <int OUT> __global__ void __launch_bounds__(kThreads) bf16_kernel(const uint16_t *__restrict__ src, void *__restrict__ dst, long long n)
{
    using O = typename OutT<OUT>::type;
    O *out = reinterpret_cast<O *>(dst);
    const long long stride = (long long)gridDim.x * kThreads;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const long long n8 = vec_ok ? n / 8 : 0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n8; i += stride) {
        uint4 w = *reinterpret_cast<const uint4 *>(src + i * 8);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        if constexpr (OUT == kBF16) {
            st_global_v4(out + i * 8, ws[0], ws[1], ws[2], ws[3]);
        } else if constexpr (OUT == kF32) {
            st_global_v4(out + i * 8, ws[0] << 16, ws[0] & 0xFFFF0000u, ws[1] << 16, ws[1] & 0xFFFF0000u);
            st_global_v4(out + i * 8 + 4, ws[2] << 16, ws[2] & 0xFFFF0000u, ws[3] << 16, ws[3] & 0xFFFF0000u);
        } else {
            uint32_t r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __half2 h = __floats2half2_rn(__uint_as_float(ws[j] << 16), __uint_as_float(ws[j] & 0xFFFF0000u));
                r[j] = *reinterpret_cast<uint32_t *>(&h);
            }
            st_global_v4(out + i * 8, r[0], r[1], r[2], r[3]);
        }
    }
    for (long long i = n8 * 8 + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float f = __uint_as_float((uint32_t)src[i] << 16);
        if constexpr (OUT == kF16) out[i] = __float2half_rn(f);
        else if constexpr (OUT == kBF16) out[i] = __float2bfloat16_rn(f);
        else out[i] = f;
    }
}

// integer-unpack debug kernel: exercises exactly the q4()/scales() the product kernels use
template <class Q>
__global__ void unpack_int_kernel(const uint8_t *__restrict__ src, long long n_blocks, int16_t *q, int16_t *sc, int16_t *mn)
{
    const long long n4 = n_blocks * (Q::BS / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (Q::BS / 4);
        const int e0 = (int)(i % (Q::BS / 4)) * 4;
        // global pointers carry no alignment guarantee here -> byte-assemble through a local copy
        uint8_t local[Q::TS + 2] __attribute__((aligned(16)));
        for (int k = 0; k < Q::TS; ++k) local[k] = src[b * Q::TS + k];
        uint32_t u = Q::q4(local, e0);
        int s, m;
        Q::scales(local, e0, s, m);
        for (int j = 0; j < 4; ++j) {
            long long o = b * Q::BS + e0 + j;
            if (q) q[o] = (int16_t)((int)((u >> (8 * j)) & 0xFF) - Q::BIAS);
            if (sc) sc[o] = (int16_t)s;
            if (mn) mn[o] = (int16_t)m;
        }
    }
}

// ------------------------------------------------------------------ host-side dispatch
template <class Q, int MATH, int OUT> static int launch_dequant(const void *packed, long long n_blocks, void *out, cudaStream_t st)
{
    constexpr int STAGES = 3;
    constexpr int TILE_ELEMS = kThreads * 32;
    constexpr int TILE_BLOCKS = TILE_ELEMS / Q::BS;
    constexpr int SLOT_BYTES = TILE_BLOCKS * Q::TS + 16;
    constexpr int SMEM = 128 + STAGES * SLOT_BYTES + 2 * TILE_ELEMS * OutT<OUT>::bytes;
    auto kern = dequant_kernel<Q, MATH, OUT, STAGES, TILE_ELEMS>;
    static unsigned char smem_set[64] = {};
    static int resident_on[64] = {};   // CTAs of this instantiation that fit on one SM, per device
    if (!ensure_dynamic_smem(kern, SMEM, smem_set)) return GGUFB200_E_CUDA;
    int &resident = resident_on[device_slot()];
    if (resident == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kThreads, SMEM) != cudaSuccess || n < 1) n = 1;
        resident = n > 4 ? 4 : n;
    }
    long long n_tiles = (n_blocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    int per_sm = g_dequant_ctas_per_sm > 0 ? g_dequant_ctas_per_sm : resident;
    if (per_sm > resident) per_sm = resident;
    long long cap = (long long)sm_count() * per_sm;
    long long rounds = (n_tiles + cap - 1) / cap;            // every CTA gets the same number of tiles (+-1)
    long long grid = (n_tiles + rounds - 1) / rounds;
    int bulk_ok = ((reinterpret_cast<uintptr_t>(packed) & 15) == 0) ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_dequant_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, reinterpret_cast<const uint8_t *>(packed), out, (long long)n_blocks, bulk_ok);
    return e == cudaSuccess ? GGUFB200_OK : GGUFB200_E_CUDA;
}

template <class Q, int MATH> static int dispatch_out(const void *packed, long long n_blocks, void *out, int out_dtype, cudaStream_t st)
{
    switch (out_dtype) {
    case kF16: return launch_dequant<Q, MATH, kF16>(packed, n_blocks, out, st);
    case kBF16: return launch_dequant<Q, MATH, kBF16>(packed, n_blocks, out, st);
    case kF32: return launch_dequant<Q, MATH, kF32>(packed, n_blocks, out, st);
    }
    return GGUFB200_E_DTYPE;
}

template <class Q> static int dispatch_math(const void *packed, long long n_blocks, void *out, int out_dtype, int math_dtype, cudaStream_t st)
{
    switch (math_dtype) {
    case kF16: return dispatch_out<Q, kF16>(packed, n_blocks, out, out_dtype, st);
    case kBF16: return dispatch_out<Q, kBF16>(packed, n_blocks, out, out_dtype, st);
    case kF32: return dispatch_out<Q, kF32>(packed, n_blocks, out, out_dtype, st);
    }
    return GGUFB200_E_DTYPE;
}

int dequant_dispatch(int type, const void *packed, long long n_blocks, void *out, int out_dtype, int math_dtype, cudaStream_t st)
{
    if (n_blocks == 0) return GGUFB200_OK;
    switch (type) {
    case T_Q4_0: return dispatch_math<Block<T_Q4_0>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q4_1: return dispatch_math<Block<T_Q4_1>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q5_0: return dispatch_math<Block<T_Q5_0>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q5_1: return dispatch_math<Block<T_Q5_1>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q8_0: return dispatch_math<Block<T_Q8_0>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q2_K: return dispatch_math<Block<T_Q2_K>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q3_K: return dispatch_math<Block<T_Q3_K>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q4_K: return dispatch_math<Block<T_Q4_K>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q5_K: return dispatch_math<Block<T_Q5_K>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_Q6_K: return dispatch_math<Block<T_Q6_K>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_IQ4_NL: return dispatch_math<Block<T_IQ4_NL>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_IQ4_XS: return dispatch_math<Block<T_IQ4_XS>>(packed, n_blocks, out, out_dtype, math_dtype, st);
    case T_BF16: {
        long long blocks = (n_blocks + (long long)kThreads * 8 - 1) / ((long long)kThreads * 8);
        long long cap = (long long)sm_count() * 8;
        unsigned grid = (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
        const uint16_t *s = reinterpret_cast<const uint16_t *>(packed);
        if (out_dtype == kF16) bf16_kernel<kF16><<<grid, kThreads, 0, st>>>(s, out, n_blocks);
        else if (out_dtype == kBF16) bf16_kernel<kBF16><<<grid, kThreads, 0, st>>>(s, out, n_blocks);
        else if (out_dtype == kF32) bf16_kernel<kF32><<<grid, kThreads, 0, st>>>(s, out, n_blocks);
        else return GGUFB200_E_DTYPE;
        return cudaGetLastError() == cudaSuccess ? GGUFB200_OK : GGUFB200_E_CUDA;
    }
    }
    return GGUFB200_E_TYPE;
}

template <class Q> static int launch_unpack(const void *packed, long long n_blocks, int16_t *q, int16_t *sc, int16_t *mn, cudaStream_t st)
{
    long long n4 = n_blocks * (Q::BS / 4);
    long long blocks = (n4 + 127) / 128;
    if (blocks > 65535) blocks = 65535;
    unpack_int_kernel<Q><<<(unsigned)blocks, 128, 0, st>>>(reinterpret_cast<const uint8_t *>(packed), n_blocks, q, sc, mn);
    return cudaGetLastError() == cudaSuccess ? GGUFB200_OK : GGUFB200_E_CUDA;
}

int unpack_dispatch(int type, const void *packed, long long n_blocks, int16_t *q, int16_t *sc, int16_t *mn, cudaStream_t st)
{
    if (n_blocks == 0) return GGUFB200_OK;
    switch (type) {
    case T_Q4_0: return launch_unpack<Block<T_Q4_0>>(packed, n_blocks, q, sc, mn, st);
    case T_Q4_1: return launch_unpack<Block<T_Q4_1>>(packed, n_blocks, q, sc, mn, st);
    case T_Q5_0: return launch_unpack<Block<T_Q5_0>>(packed, n_blocks, q, sc, mn, st);
    case T_Q5_1: return launch_unpack<Block<T_Q5_1>>(packed, n_blocks, q, sc, mn, st);
    case T_Q8_0: return launch_unpack<Block<T_Q8_0>>(packed, n_blocks, q, sc, mn, st);
    case T_Q2_K: return launch_unpack<Block<T_Q2_K>>(packed, n_blocks, q, sc, mn, st);
    case T_Q3_K: return launch_unpack<Block<T_Q3_K>>(packed, n_blocks, q, sc, mn, st);
    case T_Q4_K: return launch_unpack<Block<T_Q4_K>>(packed, n_blocks, q, sc, mn, st);
    case T_Q5_K: return launch_unpack<Block<T_Q5_K>>(packed, n_blocks, q, sc, mn, st);
    case T_Q6_K: return launch_unpack<Block<T_Q6_K>>(packed, n_blocks, q, sc, mn, st);
    case T_IQ4_NL: return launch_unpack<Block<T_IQ4_NL>>(packed, n_blocks, q, sc, mn, st);
    case T_IQ4_XS: return launch_unpack<Block<T_IQ4_XS>>(packed, n_blocks, q, sc, mn, st);
    }
    return GGUFB200_E_TYPE;
}

}  // namespace ggufb200
