// api.cu -- the extern "C" boundary declared in include/ggufb200.h.
// Argument validation lives here; kernels live in dequant.cu / rows.cu / gemv.cu / gemm.cu.
#include "blocks.cuh"

namespace ggufb200 {
extern int g_dequant_ctas_per_sm;
extern int g_dequant_pdl;
extern int g_fused_staged;
extern int g_gemv_mma;
extern int g_fused_splitk;
int dequant_dispatch(int type, const void *packed, long long n_blocks, void *out, int out_dtype, int math_dtype, cudaStream_t st);
int unpack_dispatch(int type, const void *packed, long long n_blocks, int16_t *q, int16_t *sc, int16_t *mn, cudaStream_t st);
int rows_dispatch(int type, const void *packed, long long n_table_rows, long long K, const long long *rows, long long n_rows,
                  void *out, int out_dtype, int math_dtype, cudaStream_t st);
int gemv_dispatch(int type, const void *W, long long N, long long K, const void *X, long long M, long long ldx, int act_dtype,
                  int math_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st);
int gemm_fused_dispatch(int type, const void *W, long long N, long long K, const void *X, long long M, long long ldx, int act_dtype,
                        int math_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st);
int gemm_dense_dispatch(const void *W, long long N, long long K, long long ldw, const void *X, long long M, long long ldx,
                        int act_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st);
int gemm2_fused_dispatch(int type, const void *W, long long N, long long K, const void *X, long long M, long long ldx, int act_dtype,
                         int math_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, void *ws, size_t ws_bytes, cudaStream_t st);
int gemm2_fused_splits(long long M, long long N, long long K);
void gemm2_fused_plan_info(long long M, long long N, long long K, size_t ws_bytes, int *accs, int *splits, int *kb_per_split, int *ctas);
int gemm2_dense_dispatch(const void *W, long long N, long long K, long long ldw, const void *X, long long M, long long ldx,
                         int act_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st);
int gemm3_dense_dispatch(const void *W, long long N, long long K, long long ldw, const void *X, long long M, long long ldx,
                         int act_dtype, const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st);
int gemm_fused_supported(int type);
int gemv_max_m();
}  // namespace ggufb200

using namespace ggufb200;

static int g_auto_fused = 0;     // large-M route picked by GGUFB200_ALGO_AUTO: 0 = dequant + tensor-core GEMM, 1 = fused; set_tuning(3, v)
// 0 = single-CTA UMMA (gemm.cu), 1 = CTA-pair UMMA cta_group::2 (gemm2.cu), 2 = 1 + persistent double-buffered dense GEMM (gemm3.cu)
static int g_gemm_variant = 2;   // ggufb200_set_tuning(2, v)

static int fused_mma(int type, const void *W, long long N, long long K, const void *X, long long M, long long ldx, int act, int math,
                     const void *bias, int bias_dtype, void *Y, long long ldy, void *ws, size_t ws_bytes, cudaStream_t st)
{
    return g_gemm_variant >= 1 ? gemm2_fused_dispatch(type, W, N, K, X, M, ldx, act, math, bias, bias_dtype, Y, ldy, ws, ws_bytes, st)
                               : gemm_fused_dispatch(type, W, N, K, X, M, ldx, act, math, bias, bias_dtype, Y, ldy, st);
}
// GGUFB200_ALGO_AUTO for M > 8, measured on B200 (profiles/): with M >= ~1k the dequant-once + dense GEMM route wins
// (1.27-1.44 vs 0.87-1.21 PFLOP/s); for short activations the fused kernel wins because the standalone dequant is no longer
// amortised over many M tiles (and split-K keeps all SM pairs busy while every packed byte is still read once).
static bool auto_prefers_fused(int type, long long M, long long N, long long K)
{
    if (!gemm_fused_supported(type) || (K % 64) != 0) return false;
    if (g_auto_fused) return true;
    if (M > 1024) return false;
    if (g_gemm_variant >= 1) {   // split-K: all SM pairs dequantise in parallel (measured: profiles/r01_bench_linear_graph_m512_m64.log)
        const int splits = gemm2_fused_splits(M, N, K);
        if (splits >= 4 || (splits >= 2 && K >= 2 * N)) return true;
        if (splits >= 2) return false;
    }
    return M > 256 && N >= 12288;
}

static int dense_mma(const void *W, long long N, long long K, long long ldw, const void *X, long long M, long long ldx, int act,
                     const void *bias, int bias_dtype, void *Y, long long ldy, cudaStream_t st)
{
    if (g_gemm_variant == 2) return gemm3_dense_dispatch(W, N, K, ldw, X, M, ldx, act, bias, bias_dtype, Y, ldy, st);
    return g_gemm_variant == 1 ? gemm2_dense_dispatch(W, N, K, ldw, X, M, ldx, act, bias, bias_dtype, Y, ldy, st)
                               : gemm_dense_dispatch(W, N, K, ldw, X, M, ldx, act, bias, bias_dtype, Y, ldy, st);
}

static bool type_geom(int t, int *bs, int *ts)
{
    int b = 0, s = 0;
    switch (t) {
    case T_Q4_0: b = 32; s = 18; break;
    case T_Q4_1: b = 32; s = 20; break;
    case T_Q5_0: b = 32; s = 22; break;
    case T_Q5_1: b = 32; s = 24; break;
    case T_Q8_0: b = 32; s = 34; break;
    case T_Q2_K: b = 256; s = 84; break;
    case T_Q3_K: b = 256; s = 110; break;
    case T_Q4_K: b = 256; s = 144; break;
    case T_Q5_K: b = 256; s = 176; break;
    case T_Q6_K: b = 256; s = 210; break;
    case T_IQ4_NL: b = 32; s = 18; break;
    case T_IQ4_XS: b = 256; s = 136; break;
    case T_BF16: b = 1; s = 2; break;
    default: return false;
    }
    if (bs) *bs = b;
    if (ts) *ts = s;
    return true;
}

static bool dtype_ok(int d) { return d >= 0 && d <= 2; }
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" {

int ggufb200_version(void) { return GGUFB200_VERSION; }

const char *ggufb200_strerror(int rc)
{
    switch (rc) {
    case GGUFB200_OK: return "ok";
    case GGUFB200_E_TYPE: return "unsupported ggml quantization type (no CPU fallback is provided)";
    case GGUFB200_E_DTYPE: return "dtype code must be 0 (float16), 1 (bfloat16) or 2 (float32)";
    case GGUFB200_E_ALIGN: return "output / activation pointers must be 16-byte aligned";
    case GGUFB200_E_SHAPE: return "bad shape: sizes must be non-negative, K a multiple of the block size, ld >= row length";
    case GGUFB200_E_NULL: return "required pointer is NULL";
    case GGUFB200_E_CUDA: return "CUDA launch failed";
    case GGUFB200_E_WORKSPACE: return "workspace too small (see ggufb200_linear_workspace)";
    case GGUFB200_E_UNSUPPORTED: return "operation not implemented for this type / dtype combination";
    case GGUFB200_E_DEVICE: return "current CUDA device is not sm_100 (B200)";
    }
    return "unknown error";
}

int ggufb200_type_info(int ggml_type, int *block_size, int *type_size)
{
    return type_geom(ggml_type, block_size, type_size) ? GGUFB200_OK : GGUFB200_E_TYPE;
}

int ggufb200_supported(int ggml_type, int op)
{
    if (!type_geom(ggml_type, nullptr, nullptr)) return 0;
    switch (op) {
    case GGUFB200_OP_DEQUANT: return 1;
    case GGUFB200_OP_ROWS: return 1;
    case GGUFB200_OP_LINEAR: return 1;
    case GGUFB200_OP_LINEAR_MMA: return 1;   // fused kernel, or dequant + tensor-core GEMM for types / shapes it does not cover
    }
    return 0;
}

int ggufb200_set_tuning(int key, int value)
{
    if (key == 0) {
        g_dequant_ctas_per_sm = value;
        return GGUFB200_OK;
    }
    if (key == 1) {
        g_dequant_pdl = value ? 1 : 0;
        return GGUFB200_OK;
    }
    if (key == 2) {
        g_gemm_variant = value < 0 ? 0 : (value > 2 ? 2 : value);
        return GGUFB200_OK;
    }
    if (key == 3) {
        g_auto_fused = value ? 1 : 0;
        return GGUFB200_OK;
    }
    if (key == 4) {
        g_fused_staged = value ? 1 : 0;
        return GGUFB200_OK;
    }
    if (key == 5) {
        g_gemv_mma = value ? 1 : 0;
        return GGUFB200_OK;
    }
    if (key == 6) {
        g_fused_splitk = value ? 1 : 0;
        return GGUFB200_OK;
    }
    return GGUFB200_E_UNSUPPORTED;
}

int ggufb200_dequant(int ggml_type, const void *packed, int64_t n_blocks, void *out, int out_dtype, int math_dtype, void *stream)
{
    if (!type_geom(ggml_type, nullptr, nullptr)) return GGUFB200_E_TYPE;
    if (!dtype_ok(out_dtype) || !dtype_ok(math_dtype)) return GGUFB200_E_DTYPE;
    if (n_blocks < 0) return GGUFB200_E_SHAPE;
    if (n_blocks == 0) return GGUFB200_OK;
    if (!packed || !out) return GGUFB200_E_NULL;
    if (!aligned16(out)) return GGUFB200_E_ALIGN;
    return dequant_dispatch(ggml_type, packed, n_blocks, out, out_dtype, math_dtype, (cudaStream_t)stream);
}

int ggufb200_unpack_int(int ggml_type, const void *packed, int64_t n_blocks, int16_t *q, int16_t *sc, int16_t *mn, void *stream)
{
    if (!type_geom(ggml_type, nullptr, nullptr) || ggml_type == T_BF16) return GGUFB200_E_TYPE;
    if (n_blocks < 0) return GGUFB200_E_SHAPE;
    if (n_blocks == 0) return GGUFB200_OK;
    if (!packed) return GGUFB200_E_NULL;
    return unpack_dispatch(ggml_type, packed, n_blocks, q, sc, mn, (cudaStream_t)stream);
}

int ggufb200_dequant_rows(int ggml_type, const void *packed, int64_t n_table_rows, int64_t K, const int64_t *rows, int64_t n_rows,
                          void *out, int out_dtype, int math_dtype, void *stream)
{
    int bs, ts;
    if (!type_geom(ggml_type, &bs, &ts)) return GGUFB200_E_TYPE;
    if (!dtype_ok(out_dtype) || !dtype_ok(math_dtype)) return GGUFB200_E_DTYPE;
    if (n_rows < 0 || n_table_rows < 0 || K <= 0 || K % bs != 0 || K % 8 != 0) return GGUFB200_E_SHAPE;
    if (n_rows == 0) return GGUFB200_OK;
    if (!packed || !rows || !out) return GGUFB200_E_NULL;
    if (!aligned16(out)) return GGUFB200_E_ALIGN;
    return rows_dispatch(ggml_type, packed, n_table_rows, K, (const long long *)rows, n_rows, out, out_dtype, math_dtype,
                         (cudaStream_t)stream);
}

size_t ggufb200_linear_workspace(int ggml_type, int64_t M, int64_t N, int64_t K, int act_dtype, int algo)
{
    if (!type_geom(ggml_type, nullptr, nullptr) || N <= 0 || K <= 0) return 0;
    const size_t dense = (size_t)N * (size_t)K * (act_dtype == kF32 ? 4 : 2);
    const size_t splitk = (g_gemm_variant >= 1 && gemm_fused_supported(ggml_type) && gemm2_fused_splits(M, N, K) > 1)
                              ? (size_t)gemm2_fused_splits(M, N, K) * (size_t)M * (size_t)N * 4 : 0;
    if (algo == GGUFB200_ALGO_DEQUANT_MMA) return dense;
    if (algo == GGUFB200_ALGO_FUSED_MMA) return splitk;          // fp32 partial-result slices of the split-K fused kernel
    if (algo == GGUFB200_ALGO_AUTO && M > gemv_max_m()) return auto_prefers_fused(ggml_type, M, N, K) ? splitk : dense;
    return 0;
}

int ggufb200_linear(int ggml_type, const void *W_packed, int64_t N, int64_t K, const void *X, int64_t M, int64_t ldx, int act_dtype,
                    int math_dtype, const void *bias, int bias_dtype, void *Y, int64_t ldy, void *workspace, size_t workspace_bytes,
                    int algo, void *stream)
{
    int bs, ts;
    if (!type_geom(ggml_type, &bs, &ts)) return GGUFB200_E_TYPE;
    if (act_dtype != kF16 && act_dtype != kBF16) return GGUFB200_E_DTYPE;
    if (!dtype_ok(math_dtype) || (bias && !dtype_ok(bias_dtype))) return GGUFB200_E_DTYPE;
    if (M < 0 || N <= 0 || K <= 0 || K % bs != 0 || K % 8 != 0 || ldx < K || ldy < N) return GGUFB200_E_SHAPE;
    if (M == 0) return GGUFB200_OK;
    if (!W_packed || !X || !Y) return GGUFB200_E_NULL;
    if (!aligned16(X) || !aligned16(Y) || (ldx % 8) != 0 || (ldy % 8) != 0) return GGUFB200_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;

    // The GEMV / fused producers read the packed rows with the per-format natural alignment (up to 16 bytes).  A packed
    // tensor that does not start on a 16-byte boundary (never produced by torch allocations, only by byte-offset views)
    // is therefore always routed through the standalone dequant kernel, which stages any alignment, plus the dense GEMM.
    if (!aligned16(W_packed)) {
        if (!workspace || workspace_bytes < (size_t)N * (size_t)K * 2) return GGUFB200_E_ALIGN;
        algo = GGUFB200_ALGO_DEQUANT_MMA;
    }

    if (algo == GGUFB200_ALGO_AUTO) {
        const bool ws_ok = workspace && workspace_bytes >= (size_t)N * (size_t)K * 2;
        const bool fused_ok = gemm_fused_supported(ggml_type) && math_dtype == kF16 && (K % 64) == 0;
        if (M <= gemv_max_m()) algo = GGUFB200_ALGO_GEMV;
        else if (fused_ok && (auto_prefers_fused(ggml_type, M, N, K) || !ws_ok)) algo = GGUFB200_ALGO_FUSED_MMA;
        else algo = GGUFB200_ALGO_DEQUANT_MMA;
    }
    switch (algo) {
    case GGUFB200_ALGO_GEMV:
        return gemv_dispatch(ggml_type, W_packed, N, K, X, M, ldx, act_dtype, math_dtype, bias, bias_dtype, Y, ldy, st);
    case GGUFB200_ALGO_FUSED_MMA:
        if (!gemm_fused_supported(ggml_type)) return GGUFB200_E_UNSUPPORTED;
        return fused_mma(ggml_type, W_packed, N, K, X, M, ldx, act_dtype, math_dtype, bias, bias_dtype, Y, ldy, workspace, workspace_bytes, st);
    case GGUFB200_ALGO_DEQUANT_MMA: {
        size_t need = (size_t)N * (size_t)K * 2;
        if (!workspace || workspace_bytes < need) return GGUFB200_E_WORKSPACE;
        if (!aligned16(workspace)) return GGUFB200_E_ALIGN;
        int rc = dequant_dispatch(ggml_type, W_packed, N * (K / bs), workspace, act_dtype, math_dtype, st);
        if (rc != GGUFB200_OK) return rc;
        return dense_mma(workspace, N, K, K, X, M, ldx, act_dtype, bias, bias_dtype, Y, ldy, st);
    }
    }
    return GGUFB200_E_UNSUPPORTED;
}

int ggufb200_linear_plan(int ggml_type, int64_t M, int64_t N, int64_t K, size_t workspace_bytes, int *tile_rows, int *k_ranges,
                         int *kblocks_per_range, int *ctas)
{
    if (!type_geom(ggml_type, nullptr, nullptr)) return GGUFB200_E_TYPE;
    if (!tile_rows || !k_ranges || !kblocks_per_range || !ctas) return GGUFB200_E_NULL;
    if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || N % 8 != 0) return GGUFB200_E_SHAPE;
    if (!gemm_fused_supported(ggml_type) || g_gemm_variant < 1) return GGUFB200_E_UNSUPPORTED;
    int accs = 1;
    gemm2_fused_plan_info(M, N, K, workspace_bytes, &accs, k_ranges, kblocks_per_range, ctas);
    *tile_rows = 256 * accs;
    return GGUFB200_OK;
}

int ggufb200_gemm(const void *W, int64_t N, int64_t K, int64_t ldw, const void *X, int64_t M, int64_t ldx, int act_dtype,
                  const void *bias, int bias_dtype, void *Y, int64_t ldy, void *stream)
{
    if (act_dtype != kF16 && act_dtype != kBF16) return GGUFB200_E_DTYPE;
    if (bias && !dtype_ok(bias_dtype)) return GGUFB200_E_DTYPE;
    if (M < 0 || N <= 0 || K <= 0 || K % 8 != 0 || ldw < K || ldx < K || ldy < N) return GGUFB200_E_SHAPE;
    if (M == 0) return GGUFB200_OK;
    if (!W || !X || !Y) return GGUFB200_E_NULL;
    if (!aligned16(W) || !aligned16(X) || !aligned16(Y) || (ldw % 8) || (ldx % 8) || (ldy % 8)) return GGUFB200_E_ALIGN;
    return dense_mma(W, N, K, ldw, X, M, ldx, act_dtype, bias, bias_dtype, Y, ldy, (cudaStream_t)stream);
}

}  // extern "C"
