/*
 * oracle_jacobi_adaptive.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Adaptive-precision block-Jacobi of the reference executor, restated in plain C; included from
 * oracle_impl.h once per (value type V, index type I).  Follows
 *   reference/preconditioner/jacobi_kernels.cpp:113-147 extract_block, :150-205 choose_pivot / swap_rows /
 *       apply_gauss_jordan_transform, :262-278 invert_block, :281-307
 *       validate_precision_reduction_feasibility, :313-410 generate (conditioning, per-block precision
 *       detection, ONE precision per storage group, conversion on store), :415-520 apply_block / apply /
 *       simple_apply with a stored precision, :597-627 transpose_jacobi
 *   reference/components/matrix_operations.hpp:21-36 compute_inf_norm (reads the row-major block as
 *       column-major: result = max over i of sum over j of |block[j][i]|, j ascending)
 *   core/preconditioner/jacobi_utils.hpp:100-150 get_supported_storage_reductions (the short-circuit
 *       order decides WHICH verifications run, and a verification that did not run counts as "unknown",
 *       not as "failed"), :171-189 get_optimal_storage_reduction
 * Storage types and their conversions: oracle_precision.h.
 * Pinned against the real reference (gko::preconditioner::Jacobi with storage_optimization) by
 * tests/test_jacobi_adaptive_cpu.py.
 */

#define ORC_IS_F64 (sizeof(V) == 8)

static V FNI(jac_round_trip)(V v, int kind)
{
    unsigned char slot[8];
    if (ORC_IS_F64) {
        orc_store_f64(slot, 0, kind, (double)v);
        return (V)orc_load_f64(slot, 0, kind);
    }
    orc_store_f32(slot, 0, kind, (float)v);
    return (V)orc_load_f32(slot, 0, kind);
}

/* compute_inf_norm on a row-major bs x bs block with row stride bs (see the header comment) */
static V FNI(jac_inf_norm)(int bs, const V* blk)
{
    V result = 0;
    for (int i = 0; i < bs; ++i) {
        V tmp = 0;
        for (int j = 0; j < bs; ++j) tmp += FABS(blk[i + j * bs]);
        if (tmp > result) result = tmp;
    }
    return result;
}

/* invert_block; returns 0 when a zero pivot stops it (the block is left as it is at that point) */
static int FNI(jac_invert)(int bs, int* perm, V* blk)
{
    for (int k = 0; k < bs; ++k) {
        int cp = 0;
        const V* colk = blk + k * bs + k;
        for (int i = 1; i < bs - k; ++i)
            if (FABS(colk[cp * bs]) < FABS(colk[i * bs])) cp = i;
        cp += k;
        for (int i = 0; i < bs; ++i) {
            const V t = blk[k * bs + i];
            blk[k * bs + i] = blk[cp * bs + i];
            blk[cp * bs + i] = t;
        }
        {
            const int t = perm[k];
            perm[k] = perm[cp];
            perm[cp] = t;
        }
        const V d = blk[k * bs + k];
        if (d == 0) return 0;
        for (int i = 0; i < bs; ++i) blk[i * bs + k] /= -d;
        blk[k * bs + k] = 0;
        for (int i = 0; i < bs; ++i)
            for (int j = 0; j < bs; ++j) blk[i * bs + j] += blk[i * bs + k] * blk[k * bs + j];
        for (int j = 0; j < bs; ++j) blk[k * bs + j] /= d;
        blk[k * bs + k] = (V)1 / d;
    }
    return 1;
}

/* validate_precision_reduction_feasibility<ReducedType>: the INVERTED block rounded to the reduced
 * type must itself be invertible with a condition number in [1, 1e-3 / eps(V)) */
static int FNI(jac_validate)(int bs, const V* inv, int kind)
{
    V tmp[32 * 32];
    int perm[32];
    for (int i = 0; i < bs; ++i) perm[i] = i;
    for (int i = 0; i < bs; ++i)
        for (int j = 0; j < bs; ++j) tmp[i * bs + j] = FNI(jac_round_trip)(inv[i * bs + j], kind);
    V cond = FNI(jac_inf_norm)(bs, tmp);
    if (!FNI(jac_invert)(bs, perm, tmp)) return 0;
    cond *= FNI(jac_inf_norm)(bs, tmp);
    const V eps = ORC_IS_F64 ? (V)(1.0 / 9007199254740992.0) : (V)(1.0 / 16777216.0); /* 2^-53, 2^-24 */
    return cond >= (V)1.0 && cond * eps < (V)1e-3;
}

/* jacobi::generate with conditioning / block_precisions (both may be NULL: full precision) */
void FNI(jacobi_generate_adaptive)(int64_t num_rows, const I* rp, const I* ci, const V* va, int64_t num_blocks,
                                   int32_t max_block_size, double accuracy_d, int64_t block_offset,
                                   int64_t group_offset, int32_t group_power, V* conditioning,
                                   uint8_t* block_precisions, const I* block_ptrs, V* blocks)
{
    (void)num_rows;
    (void)max_block_size;
    const V accuracy = (V)accuracy_d;
    const int64_t stride = block_offset << group_power;
    const int64_t group_size = (int64_t)1 << group_power;
    /* float_traits<...>::eps of the candidate storage types, core/base/extended_float.hpp + half.hpp:121-143:
     * 1 / 2^(significand bits + rounds_to_nearest) */
    const V eps_tt = ORC_IS_F64 ? (V)(1.0 / 16) : (V)(1.0 / 128);         /* truncate(truncate(V)) */
    const V eps_tr = ORC_IS_F64 ? (V)(1.0 / 128) : (V)(1.0 / 2048);       /* truncate(reduce(V)) */
    const V eps_rr = (V)(1.0 / 2048);                                      /* reduce(reduce(V)) = half */
    const V eps_t = ORC_IS_F64 ? (V)(1.0 / 1048576) : (V)(1.0 / 128);     /* truncate(V) */
    const V eps_r = ORC_IS_F64 ? (V)(1.0 / 16777216.0) : (V)(1.0 / 2048); /* reduce(V) */
    const int kind_r = ORC_IS_F64 ? ORC_ST_F32 : ORC_ST_F16;
    const int kind_rr = ORC_ST_F16;
    V* blk = (V*)malloc(sizeof(V) * 32 * 32 * (size_t)group_size);
    int* perm = (int*)malloc(sizeof(int) * 32 * (size_t)group_size);
    for (int64_t g = 0; g < num_blocks; g += group_size) {
        uint32_t descr = 0xffffffffu;
        for (int64_t bq = 0; bq < group_size && g + bq < num_blocks; ++bq) {
            V* bl = blk + 1024 * bq;
            int* pm = perm + 32 * bq;
            const int64_t start = block_ptrs[g + bq];
            const int bs = (int)((int64_t)block_ptrs[g + bq + 1] - start);
            for (int i = 0; i < bs * bs; ++i) bl[i] = 0;
            for (int i = 0; i < bs; ++i) pm[i] = i;
            for (int row = 0; row < bs; ++row)
                for (int64_t p = rp[start + row]; p < (int64_t)rp[start + row + 1]; ++p) {
                    const int64_t col = (int64_t)ci[p] - start;
                    if (0 <= col && col < bs) bl[row * bs + col] = va[p];
                }
            if (conditioning) conditioning[g + bq] = FNI(jac_inf_norm)(bs, bl);
            FNI(jac_invert)(bs, pm, bl);
            if (conditioning) conditioning[g + bq] *= FNI(jac_inf_norm)(bs, bl);
            const uint8_t local = block_precisions ? block_precisions[g + bq] : 0;
            uint32_t d;
            if (local == 0xff && conditioning) {
                const V cond = conditioning[g + bq];
                int v1 = 2; /* 2: not evaluated */
                d = ORC_P0N0;
                if (cond * eps_tt < accuracy) d |= ORC_P2N0;
                if (cond * eps_tr < accuracy && (v1 = FNI(jac_validate)(bs, bl, kind_r))) d |= ORC_P1N1;
                if (cond * eps_rr < accuracy && v1 != 0 && FNI(jac_validate)(bs, bl, kind_rr)) d |= ORC_P0N2;
                if (cond * eps_t < accuracy) d |= ORC_P1N0;
                if (cond * eps_r < accuracy &&
                    (v1 == 1 || (v1 == 2 && (v1 = FNI(jac_validate)(bs, bl, kind_r)))))
                    d |= ORC_P0N1;
            } else {
                d = orc_prd_singleton(local);
            }
            descr &= d;
        }
        const uint8_t p = orc_optimal_reduction(descr);
        const int kind = orc_storage_kind(ORC_IS_F64, p);
        for (int64_t bq = 0; bq < group_size && g + bq < num_blocks; ++bq) {
            const V* bl = blk + 1024 * bq;
            const int* pm = perm + 32 * bq;
            const int64_t k = g + bq;
            if (block_precisions) block_precisions[k] = p;
            const int bs = (int)((int64_t)block_ptrs[k + 1] - (int64_t)block_ptrs[k]);
            void* group = (void*)(blocks + group_offset * (k >> group_power));
            const int64_t bo = block_offset * (k & (group_size - 1));
            for (int i = 0; i < bs; ++i)
                for (int j = 0; j < bs; ++j) {
                    const int64_t idx = bo + i + pm[j] * stride;
                    if (ORC_IS_F64)
                        orc_store_f64(group, idx, kind, (double)bl[i * bs + j]);
                    else
                        orc_store_f32(group, idx, kind, (float)bl[i * bs + j]);
                }
        }
    }
    free(blk);
    free(perm);
}

/* jacobi::apply with block_precisions (NULL: full precision), alpha / beta NULL: 1 / 0 */
void FNI(jacobi_apply_adaptive)(int64_t num_blocks, int32_t max_block_size, int64_t block_offset,
                                int64_t group_offset, int32_t group_power, const uint8_t* block_precisions,
                                const I* block_ptrs, const V* blocks, const V* alpha_p, const V* b, int64_t bs,
                                int64_t num_rhs, const V* beta_p, V* x, int64_t xs)
{
    (void)max_block_size;
    const V alpha = alpha_p ? alpha_p[0] : (V)1;
    const V beta = beta_p ? beta_p[0] : (V)0;
    const int64_t stride = block_offset << group_power;
    for (int64_t k = 0; k < num_blocks; ++k) {
        const void* group = (const void*)(blocks + group_offset * (k >> group_power));
        const int64_t bo = block_offset * (k & (((int64_t)1 << group_power) - 1));
        const int kind = orc_storage_kind(ORC_IS_F64, block_precisions ? block_precisions[k] : 0);
        const int64_t first = block_ptrs[k];
        const int64_t n = (int64_t)block_ptrs[k + 1] - first;
        V* xb = x + first * xs;
        const V* bb = b + first * bs;
        for (int64_t row = 0; row < n; ++row)
            for (int64_t col = 0; col < num_rhs; ++col) {
                if (beta != 0)
                    xb[row * xs + col] *= beta;
                else
                    xb[row * xs + col] = 0;
            }
        for (int64_t inner = 0; inner < n; ++inner)
            for (int64_t row = 0; row < n; ++row) {
                const int64_t idx = bo + row + inner * stride;
                const V e = ORC_IS_F64 ? (V)orc_load_f64(group, idx, kind) : (V)orc_load_f32(group, idx, kind);
                for (int64_t col = 0; col < num_rhs; ++col) xb[row * xs + col] += alpha * e * bb[inner * bs + col];
            }
    }
}

void FNI(jacobi_simple_apply_adaptive)(int64_t num_blocks, int32_t max_block_size, int64_t block_offset,
                                       int64_t group_offset, int32_t group_power,
                                       const uint8_t* block_precisions, const I* block_ptrs, const V* blocks,
                                       const V* b, int64_t bs, int64_t num_rhs, V* x, int64_t xs)
{
    FNI(jacobi_apply_adaptive)(num_blocks, max_block_size, block_offset, group_offset, group_power,
                               block_precisions, block_ptrs, blocks, NULL, b, bs, num_rhs, NULL, x, xs);
}

/* transpose_jacobi with stored precisions: out(j, i) = in(i, j) inside every block, bits unchanged */
void FNI(jacobi_transpose_adaptive)(int64_t num_blocks, int32_t max_block_size, int64_t block_offset,
                                    int64_t group_offset, int32_t group_power, const uint8_t* block_precisions,
                                    const I* block_ptrs, const V* blocks, V* out_blocks)
{
    (void)max_block_size;
    const int64_t stride = block_offset << group_power;
    for (int64_t k = 0; k < num_blocks; ++k) {
        const char* group = (const char*)(blocks + group_offset * (k >> group_power));
        char* ogroup = (char*)(out_blocks + group_offset * (k >> group_power));
        const int64_t bo = block_offset * (k & (((int64_t)1 << group_power) - 1));
        const int kind = orc_storage_kind(ORC_IS_F64, block_precisions ? block_precisions[k] : 0);
        const int w = orc_storage_bytes(kind);
        const int64_t n = (int64_t)block_ptrs[k + 1] - (int64_t)block_ptrs[k];
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j)
                memcpy(ogroup + (bo + i * stride + j) * w, group + (bo + i + j * stride) * w, (size_t)w);
    }
}

#undef ORC_IS_F64
