/*
 * oracle_impl.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the reference executor's arithmetic for the SpMV + Krylov
 * hot path, in plain C, one function per reference kernel, each citing the
 * reference file:line it follows (paths relative to /root/reference).  Included
 * once per value type from oracle.c with
 *     #define V double / float,  #define VS f64 / f32
 * and, for the index-typed part, I / IS.  Parity is PINNED: tests/
 * test_oracle_golden.py checks every function against the literal known-answer
 * vectors of the reference's own reference/test/ suites, and
 * tests/test_oracle_vs_ref.py against the reference itself (oracle/_ref).
 *
 * Compiled with -ffp-contract=off: the reference build (x86-64 baseline) has no
 * FMA contraction, every product is rounded before it is added.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(CAT(orc, name), VS)
#define FNI(name) CAT(CAT(CAT(orc, name), VS), IS)

#ifndef ORC_VALUE_PART_DONE
/* ===================== value-typed kernels (no index type) ================= */

/* reference/matrix/dense_kernels.cpp:262-279 (compute_dot), :295-311 (conj_dot) */
void FN(dense_compute_dot)(int64_t rows, int64_t cols, const V* x, int64_t xs, const V* y,
                           int64_t ys, V* result)
{
    for (int64_t j = 0; j < cols; ++j) result[j] = 0;
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) result[j] += x[i * xs + j] * y[i * ys + j];
}

void FN(dense_compute_conj_dot)(int64_t rows, int64_t cols, const V* x, int64_t xs, const V* y,
                                int64_t ys, V* result)
{
    FN(dense_compute_dot)(rows, cols, x, xs, y, ys, result); /* real types: conj == identity */
}

/* reference/matrix/dense_kernels.cpp:328-346 (norm2), :420-435 (squared_norm2) */
void FN(dense_compute_squared_norm2)(int64_t rows, int64_t cols, const V* x, int64_t xs, V* result)
{
    for (int64_t j = 0; j < cols; ++j) result[j] = 0;
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) result[j] += x[i * xs + j] * x[i * xs + j];
}
void FN(dense_compute_norm2)(int64_t rows, int64_t cols, const V* x, int64_t xs, V* result)
{
    FN(dense_compute_squared_norm2)(rows, cols, x, xs, result);
    for (int64_t j = 0; j < cols; ++j) result[j] = SQRT(result[j]);
}

/* reference/matrix/dense_kernels.cpp:127-150 */
void FN(dense_scale)(int64_t rows, int64_t cols, const V* alpha, int64_t alpha_cols, V* x,
                     int64_t xs)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (alpha_cols == 1) {
                if (alpha[0] == 0)
                    x[i * xs + j] = 0;
                else
                    x[i * xs + j] *= alpha[0];
            } else {
                x[i * xs + j] *= alpha[j];
            }
        }
}
/* reference/matrix/dense_kernels.cpp:153-172 */
void FN(dense_inv_scale)(int64_t rows, int64_t cols, const V* alpha, int64_t alpha_cols, V* x,
                         int64_t xs)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) x[i * xs + j] /= alpha[alpha_cols == 1 ? 0 : j];
}
/* reference/matrix/dense_kernels.cpp:177-198 */
void FN(dense_add_scaled)(int64_t rows, int64_t cols, const V* alpha, int64_t alpha_cols,
                          const V* x, int64_t xs, V* y, int64_t ys)
{
    if (alpha_cols == 1 && alpha[0] == 0) return;
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j)
            y[i * ys + j] += alpha[alpha_cols == 1 ? 0 : j] * x[i * xs + j];
}
/* reference/matrix/dense_kernels.cpp:203-224 */
void FN(dense_sub_scaled)(int64_t rows, int64_t cols, const V* alpha, int64_t alpha_cols,
                          const V* x, int64_t xs, V* y, int64_t ys)
{
    if (alpha_cols == 1 && alpha[0] == 0) return;
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j)
            y[i * ys + j] -= alpha[alpha_cols == 1 ? 0 : j] * x[i * xs + j];
}

/* ---- stopping_status helpers: include/ginkgo/core/stop/stopping_status.hpp ---- */
#ifndef ORC_STATUS_HELPERS
#define ORC_STATUS_HELPERS
static int st_has_stopped(uint8_t s) { return (s & 0x3f) != 0; }
static int st_is_finalized(uint8_t s) { return (s & 0x40) != 0; }
static uint8_t st_converge(uint8_t s, uint8_t id, int set_finalized)
{
    if (!st_has_stopped(s)) {
        s |= 0x80 | (id & 0x3f);
        if (set_finalized) s |= 0x40;
    }
    return s;
}
static uint8_t st_stop(uint8_t s, uint8_t id, int set_finalized)
{
    if (!st_has_stopped(s)) {
        s |= (id & 0x3f);
        if (set_finalized) s |= 0x40;
    }
    return s;
}
static uint8_t st_finalize(uint8_t s)
{
    if (st_has_stopped(s)) s |= 0x40;
    return s;
}
/* reference/stop/criterion_kernels.cpp (set_all_statuses): stop(id, set_finalized) */
void orc_set_all_statuses(int64_t cols, uint8_t id, int32_t set_finalized, uint8_t* stop)
{
    for (int64_t i = 0; i < cols; ++i) stop[i] = st_stop(stop[i], id, set_finalized);
}

#endif

/* reference/solver/cg_kernels.cpp:24-45 */
void FN(cg_initialize)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* r, int64_t rs, V* z,
                       int64_t zs, V* p, int64_t ps, V* q, int64_t qs, V* prev_rho, V* rho,
                       uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        rho[j] = 0;
        prev_rho[j] = 1;
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            r[i * rs + j] = b[i * bs + j];
            z[i * zs + j] = p[i * ps + j] = q[i * qs + j] = 0;
        }
}
/* reference/solver/cg_kernels.cpp:50-72 */
void FN(cg_step_1)(int64_t rows, int64_t cols, V* p, int64_t ps, const V* z, int64_t zs,
                   const V* rho, const V* prev_rho, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (prev_rho[j] == 0) {
                p[i * ps + j] = z[i * zs + j];
            } else {
                V tmp = rho[j] / prev_rho[j];
                p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
            }
        }
}
/* reference/solver/cg_kernels.cpp:77-100 */
void FN(cg_step_2)(int64_t rows, int64_t cols, V* x, int64_t xs, V* r, int64_t rs, const V* p,
                   int64_t ps, const V* q, int64_t qs, const V* beta, const V* rho,
                   const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (beta[j] != 0) {
                V tmp = rho[j] / beta[j];
                x[i * xs + j] += tmp * p[i * ps + j];
                r[i * rs + j] -= tmp * q[i * qs + j];
            }
        }
}

/* reference/solver/fcg_kernels.cpp:22-100 */
void FN(fcg_initialize)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* r, int64_t rs, V* z,
                        int64_t zs, V* p, int64_t ps, V* q, int64_t qs, V* t, int64_t ts,
                        V* prev_rho, V* rho, V* rho_t, uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        rho[j] = 0;
        prev_rho[j] = 1;
        rho_t[j] = 1;
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            t[i * ts + j] = r[i * rs + j] = b[i * bs + j];
            z[i * zs + j] = p[i * ps + j] = q[i * qs + j] = 0;
        }
}
void FN(fcg_step_1)(int64_t rows, int64_t cols, V* p, int64_t ps, const V* z, int64_t zs,
                    const V* rho_t, const V* prev_rho, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (prev_rho[j] == 0) {
                p[i * ps + j] = z[i * zs + j];
            } else {
                V tmp = rho_t[j] / prev_rho[j];
                p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
            }
        }
}
void FN(fcg_step_2)(int64_t rows, int64_t cols, V* x, int64_t xs, V* r, int64_t rs, V* t,
                    int64_t ts, const V* p, int64_t ps, const V* q, int64_t qs, const V* beta,
                    const V* rho, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (beta[j] != 0) {
                V tmp = rho[j] / beta[j];
                V prev_r = r[i * rs + j];
                x[i * xs + j] += tmp * p[i * ps + j];
                r[i * rs + j] -= tmp * q[i * qs + j];
                t[i * ts + j] = r[i * rs + j] - prev_r;
            }
        }
}

/* reference/solver/cgs_kernels.cpp:22-140 */
void FN(cgs_initialize)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* r, int64_t rs,
                        V* r_tld, int64_t rts, V* p, int64_t ps, V* q, int64_t qs, V* u,
                        int64_t us, V* u_hat, int64_t uhs, V* v_hat, int64_t vhs, V* t, int64_t ts,
                        V* alpha, V* beta, V* gamma, V* prev_rho, V* rho, uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        rho[j] = 0;
        prev_rho[j] = alpha[j] = beta[j] = gamma[j] = 1;
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            r[i * rs + j] = b[i * bs + j];
            r_tld[i * rts + j] = b[i * bs + j];
            u[i * us + j] = u_hat[i * uhs + j] = p[i * ps + j] = q[i * qs + j] =
                v_hat[i * vhs + j] = t[i * ts + j] = 0;
        }
}
void FN(cgs_step_1)(int64_t rows, int64_t cols, const V* r, int64_t rs, V* u, int64_t us, V* p,
                    int64_t ps, const V* q, int64_t qs, V* beta, const V* rho, const V* prev_rho,
                    const uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        if (st_has_stopped(stop[j])) continue;
        if (prev_rho[j] != 0) beta[j] = rho[j] / prev_rho[j];
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            u[i * us + j] = r[i * rs + j] + beta[j] * q[i * qs + j];
            p[i * ps + j] = u[i * us + j] + beta[j] * (q[i * qs + j] + beta[j] * p[i * ps + j]);
        }
}
void FN(cgs_step_2)(int64_t rows, int64_t cols, const V* u, int64_t us, const V* v_hat, int64_t vhs,
                    V* q, int64_t qs, V* t, int64_t ts, V* alpha, const V* rho, const V* gamma,
                    const uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        if (st_has_stopped(stop[j])) continue;
        if (gamma[j] != 0) alpha[j] = rho[j] / gamma[j];
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            q[i * qs + j] = u[i * us + j] - alpha[j] * v_hat[i * vhs + j];
            t[i * ts + j] = u[i * us + j] + q[i * qs + j];
        }
}
void FN(cgs_step_3)(int64_t rows, int64_t cols, const V* t, int64_t ts, const V* u_hat, int64_t uhs,
                    V* r, int64_t rs, V* x, int64_t xs, const V* alpha, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            x[i * xs + j] += alpha[j] * u_hat[i * uhs + j];
            r[i * rs + j] -= alpha[j] * t[i * ts + j];
        }
}

/* reference/matrix/dense_kernels.cpp:440-448 */
void FN(dense_compute_sqrt)(int64_t rows, int64_t cols, V* data, int64_t stride)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) data[i * stride + j] = SQRT(data[i * stride + j]);
}

/* reference/solver/bicg_kernels.cpp:25-118 */
void FN(bicg_initialize)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* r, int64_t rs, V* z,
                         int64_t zs, V* p, int64_t ps, V* q, int64_t qs, V* prev_rho, V* rho, V* r2,
                         int64_t r2s, V* z2, int64_t z2s, V* p2, int64_t p2s, V* q2, int64_t q2s,
                         uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        rho[j] = 0;
        prev_rho[j] = 1;
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            r[i * rs + j] = b[i * bs + j];
            r2[i * r2s + j] = b[i * bs + j];
            z[i * zs + j] = p[i * ps + j] = q[i * qs + j] = 0;
            z2[i * z2s + j] = p2[i * p2s + j] = q2[i * q2s + j] = 0;
        }
}
void FN(bicg_step_1)(int64_t rows, int64_t cols, V* p, int64_t ps, const V* z, int64_t zs, V* p2,
                     int64_t p2s, const V* z2, int64_t z2s, const V* rho, const V* prev_rho,
                     const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (prev_rho[j] == 0) {
                p[i * ps + j] = z[i * zs + j];
                p2[i * p2s + j] = z2[i * z2s + j];
            } else {
                const V tmp = rho[j] / prev_rho[j];
                p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
                p2[i * p2s + j] = z2[i * z2s + j] + tmp * p2[i * p2s + j];
            }
        }
}
void FN(bicg_step_2)(int64_t rows, int64_t cols, V* x, int64_t xs, V* r, int64_t rs, V* r2,
                     int64_t r2s, const V* p, int64_t ps, const V* q, int64_t qs, const V* q2,
                     int64_t q2s, const V* beta, const V* rho, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (beta[j] != 0) {
                const V tmp = rho[j] / beta[j];
                x[i * xs + j] += tmp * p[i * ps + j];
                r[i * rs + j] -= tmp * q[i * qs + j];
                r2[i * r2s + j] -= tmp * q2[i * q2s + j];
            }
        }
}

/* reference/solver/ir_kernels.cpp:17-24: reset every stopping status */
#ifndef ORC_IR_INITIALIZE
#define ORC_IR_INITIALIZE
void orc_ir_initialize(int64_t cols, uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) stop[j] = 0;
}
#endif
/* reference/solver/chebyshev_kernels.cpp:15-66; coefficients and arithmetic in double
 * (solver::detail::coeff_type, include/ginkgo/core/solver/chebyshev.hpp:29-31) */
void FN(chebyshev_init_update)(int64_t rows, int64_t cols, double alpha, const V* inner_sol, int64_t is,
                               V* update_sol, int64_t us, V* output, int64_t os)
{
    for (int64_t row = 0; row < rows; ++row)
        for (int64_t col = 0; col < cols; ++col) {
            const double inner_val = (double)inner_sol[row * is + col];
            update_sol[row * us + col] = (V)inner_val;
            output[row * os + col] = (V)((double)output[row * os + col] + alpha * inner_val);
        }
}
void FN(chebyshev_update)(int64_t rows, int64_t cols, double alpha, double beta, V* inner_sol, int64_t is,
                          V* update_sol, int64_t us, V* output, int64_t os)
{
    for (int64_t row = 0; row < rows; ++row)
        for (int64_t col = 0; col < cols; ++col) {
            const double val = (double)inner_sol[row * is + col] + beta * (double)update_sol[row * us + col];
            inner_sol[row * is + col] = (V)val;
            update_sol[row * us + col] = (V)val;
            output[row * os + col] = (V)((double)output[row * os + col] + alpha * val);
        }
}

/* reference/solver/pipe_cg_kernels.cpp:24-160 */
void FN(pipe_cg_initialize_1)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* r, int64_t rs,
                              V* prev_rho, uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        prev_rho[j] = 1;
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) r[i * rs + j] = b[i * bs + j];
}
void FN(pipe_cg_initialize_2)(int64_t rows, int64_t cols, V* p, int64_t ps, V* q, int64_t qs, V* f,
                              int64_t fs, V* g, int64_t gs, V* beta, const V* z, int64_t zs,
                              const V* w, int64_t ws, const V* m, int64_t ms, const V* n, int64_t ns,
                              const V* delta)
{
    for (int64_t j = 0; j < cols; ++j) beta[j] = delta[j];
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            p[i * ps + j] = z[i * zs + j];
            q[i * qs + j] = w[i * ws + j];
            f[i * fs + j] = m[i * ms + j];
            g[i * gs + j] = n[i * ns + j];
        }
}
void FN(pipe_cg_step_1)(int64_t rows, int64_t cols, V* x, int64_t xs, V* r, int64_t rs, V* z1,
                        int64_t z1s, V* z2, int64_t z2s, V* w, int64_t ws, const V* p, int64_t ps,
                        const V* q, int64_t qs, const V* f, int64_t fs, const V* g, int64_t gs,
                        const V* rho, const V* beta, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (beta[j] != 0) {
                V tmp = rho[j] / beta[j];
                x[i * xs + j] += tmp * p[i * ps + j];
                r[i * rs + j] -= tmp * q[i * qs + j];
                z1[i * z1s + j] -= tmp * f[i * fs + j];
                z2[i * z2s + j] = z1[i * z1s + j];
                w[i * ws + j] -= tmp * g[i * gs + j];
            }
        }
}
void FN(pipe_cg_step_2)(int64_t rows, int64_t cols, V* beta, V* p, int64_t ps, V* q, int64_t qs, V* f,
                        int64_t fs, V* g, int64_t gs, const V* z, int64_t zs, const V* w, int64_t ws,
                        const V* m, int64_t ms, const V* n, int64_t ns, const V* prev_rho,
                        const V* rho, const V* delta, const uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        if (st_has_stopped(stop[j])) continue;
        if (prev_rho[j] != 0) {
            V tmp = rho[j] / prev_rho[j];
            V abs_tmp = FABS(tmp);
            beta[j] = delta[j] - abs_tmp * abs_tmp * beta[j];
            if (beta[j] == 0) beta[j] = delta[j];
            for (int64_t i = 0; i < rows; ++i) {
                p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
                q[i * qs + j] = w[i * ws + j] + tmp * q[i * qs + j];
                f[i * fs + j] = m[i * ms + j] + tmp * f[i * fs + j];
                g[i * gs + j] = n[i * ns + j] + tmp * g[i * gs + j];
            }
        } else {
            beta[j] = delta[j];
            for (int64_t i = 0; i < rows; ++i) {
                p[i * ps + j] = z[i * zs + j];
                q[i * qs + j] = w[i * ws + j];
                f[i * fs + j] = m[i * ms + j];
                g[i * gs + j] = n[i * ns + j];
            }
        }
    }
}

/* reference/solver/gcr_kernels.cpp:26-84 */
void FN(gcr_initialize)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* residual, int64_t rs,
                        uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        for (int64_t i = 0; i < rows; ++i) residual[i * rs + j] = b[i * bs + j];
        stop[j] = 0;
    }
}
void FN(gcr_restart)(int64_t rows, int64_t cols, const V* residual, int64_t rs, const V* a_residual,
                     int64_t ars, V* p_bases, int64_t ps, V* ap_bases, int64_t aps,
                     uint64_t* final_iter_nums)
{
    for (int64_t j = 0; j < cols; ++j) {
        for (int64_t i = 0; i < rows; ++i) {
            p_bases[i * ps + j] = residual[i * rs + j];
            ap_bases[i * aps + j] = a_residual[i * ars + j];
        }
        final_iter_nums[j] = 0;
    }
}
void FN(gcr_step_1)(int64_t rows, int64_t cols, V* x, int64_t xs, V* residual, int64_t rs, const V* p,
                    int64_t ps, const V* ap, int64_t aps, const V* ap_norm, const V* rap,
                    const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (ap_norm[j] != 0) {
                V tmp = rap[j] / ap_norm[j];
                x[i * xs + j] += tmp * p[i * ps + j];
                residual[i * rs + j] -= tmp * ap[i * aps + j];
            }
        }
}

/* reference/solver/minres_kernels.cpp:26-150 (safe_divide: include/ginkgo/core/base/math.hpp) */
static V FN(safe_divide)(V a, V b) { return b == 0 ? (V)0 : a / b; }
void FN(minres_initialize)(int64_t rows, int64_t cols, const V* r, int64_t rs, V* z, int64_t zs, V* p,
                           int64_t ps, V* p_prev, int64_t pps, V* q, int64_t qs, V* q_prev,
                           int64_t qps, V* q_tilde, int64_t qts, V* beta, V* gamma, V* delta,
                           V* cos_prev, V* cos_, V* sin_prev, V* sin_, V* eta_next, V* eta,
                           uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        delta[j] = gamma[j] = cos_prev[j] = sin_prev[j] = sin_[j] = 0;
        cos_[j] = 1;
        eta_next[j] = eta[j] = beta[j] = SQRT(beta[j]);
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            q[i * qs + j] = FN(safe_divide)(r[i * rs + j], beta[j]);
            z[i * zs + j] = FN(safe_divide)(z[i * zs + j], beta[j]);
            p[i * ps + j] = p_prev[i * pps + j] = q_prev[i * qps + j] = q_tilde[i * qts + j] = 0;
        }
}
void FN(minres_step_1)(int64_t cols, V* alpha, V* beta, V* gamma, V* delta, V* cos_prev, V* cos_,
                       V* sin_prev, V* sin_, V* eta, V* eta_next, V* tau, const uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        if (st_has_stopped(stop[j])) continue;
        beta[j] = SQRT(beta[j]);
        delta[j] = sin_prev[j] * gamma[j];
        const V tmp_d = gamma[j];
        const V tmp_a = alpha[j];
        gamma[j] = cos_prev[j] * cos_[j] * tmp_d + sin_[j] * tmp_a;
        alpha[j] = -sin_[j] * cos_prev[j] * tmp_d + cos_[j] * tmp_a;
        V t = cos_[j];
        cos_[j] = cos_prev[j];
        cos_prev[j] = t;
        t = sin_[j];
        sin_[j] = sin_prev[j];
        sin_prev[j] = t;
        /* update_givens_rotation(alpha, beta, cos, sin) */
        if (alpha[j] == 0) {
            cos_[j] = 0;
            sin_[j] = 1;
        } else {
            const V scale = FABS(alpha[j]) + FABS(beta[j]);
            const V hyp = scale * SQRT(FABS(alpha[j] / scale) * FABS(alpha[j] / scale) +
                                       FABS(beta[j] / scale) * FABS(beta[j] / scale));
            cos_[j] = alpha[j] / hyp;
            sin_[j] = beta[j] / hyp;
        }
        alpha[j] = cos_[j] * alpha[j] + sin_[j] * beta[j];
        tau[j] = sin_[j] * sin_[j] * tau[j];
        eta[j] = eta_next[j];
        eta_next[j] = -sin_[j] * eta[j];
    }
}
void FN(minres_step_2)(int64_t rows, int64_t cols, V* x, int64_t xs, V* p, int64_t ps, const V* p_prev,
                       int64_t pps, V* z, int64_t zs, const V* z_tilde, int64_t zts, V* q, int64_t qs,
                       V* q_prev, int64_t qps, V* v, int64_t vs, const V* alpha, const V* beta,
                       const V* gamma, const V* delta, const V* cos_, const V* eta, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            p[i * ps + j] = FN(safe_divide)(
                z[i * zs + j] - gamma[j] * p_prev[i * pps + j] - delta[j] * p[i * ps + j], alpha[j]);
            x[i * xs + j] = x[i * xs + j] + cos_[j] * eta[j] * p[i * ps + j];
            q_prev[i * qps + j] = v[i * vs + j];
            const V tmp = q[i * qs + j];
            q[i * qs + j] = FN(safe_divide)(v[i * vs + j], beta[j]);
            v[i * vs + j] = tmp * beta[j];
            z[i * zs + j] = FN(safe_divide)(z_tilde[i * zts + j], beta[j]);
        }
}

/* reference/solver/bicgstab_kernels.cpp:25-60 */
void FN(bicgstab_initialize)(int64_t rows, int64_t cols, const V* b, int64_t bs, V* r, int64_t rs,
                             V* rr, int64_t rrs, V* y, int64_t ys, V* s, int64_t ss, V* t,
                             int64_t ts, V* z, int64_t zs, V* v, int64_t vs, V* p, int64_t ps,
                             V* prev_rho, V* rho, V* alpha, V* beta, V* gamma, V* omega,
                             uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        rho[j] = prev_rho[j] = alpha[j] = beta[j] = gamma[j] = omega[j] = 1;
        stop[j] = 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            r[i * rs + j] = b[i * bs + j];
            rr[i * rrs + j] = z[i * zs + j] = v[i * vs + j] = s[i * ss + j] = t[i * ts + j] =
                y[i * ys + j] = p[i * ps + j] = 0;
        }
}
/* reference/solver/bicgstab_kernels.cpp:65-92 */
void FN(bicgstab_step_1)(int64_t rows, int64_t cols, const V* r, int64_t rs, V* p, int64_t ps,
                         const V* v, int64_t vs, const V* rho, const V* prev_rho, const V* alpha,
                         const V* omega, const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (prev_rho[j] * omega[j] != 0) {
                V tmp = rho[j] / prev_rho[j] * alpha[j] / omega[j];
                p[i * ps + j] = r[i * rs + j] + tmp * (p[i * ps + j] - omega[j] * v[i * vs + j]);
            } else {
                p[i * ps + j] = r[i * rs + j];
            }
        }
}
/* reference/solver/bicgstab_kernels.cpp:97-121 */
void FN(bicgstab_step_2)(int64_t rows, int64_t cols, const V* r, int64_t rs, V* s, int64_t ss,
                         const V* v, int64_t vs, const V* rho, V* alpha, const V* beta,
                         const uint8_t* stop)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            if (beta[j] != 0) {
                alpha[j] = rho[j] / beta[j];
                s[i * ss + j] = r[i * rs + j] - alpha[j] * v[i * vs + j];
            } else {
                alpha[j] = 0;
                s[i * ss + j] = r[i * rs + j];
            }
        }
}
/* reference/solver/bicgstab_kernels.cpp:126-158 */
void FN(bicgstab_step_3)(int64_t rows, int64_t cols, V* x, int64_t xs, V* r, int64_t rs,
                         const V* s, int64_t ss, const V* t, int64_t ts, const V* y, int64_t ys,
                         const V* z, int64_t zs, const V* alpha, const V* beta, const V* gamma,
                         V* omega, const uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        if (st_has_stopped(stop[j])) continue;
        omega[j] = beta[j] != 0 ? gamma[j] / beta[j] : 0;
    }
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            if (st_has_stopped(stop[j])) continue;
            x[i * xs + j] += alpha[j] * y[i * ys + j] + omega[j] * z[i * zs + j];
            r[i * rs + j] = s[i * ss + j] - omega[j] * t[i * ts + j];
        }
}
/* reference/solver/bicgstab_kernels.cpp:163-178 */
void FN(bicgstab_finalize)(int64_t rows, int64_t cols, V* x, int64_t xs, const V* y, int64_t ys,
                           const V* alpha, uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j)
        if (st_has_stopped(stop[j]) && !st_is_finalized(stop[j]))
            for (int64_t i = 0; i < rows; ++i) {
                x[i * xs + j] += alpha[j] * y[i * ys + j];
                stop[j] = st_finalize(stop[j]);
            }
}

/* reference/solver/common_gmres_kernels.cpp:112-133 */
void FN(common_gmres_initialize)(int64_t rows, int64_t cols, int64_t krylov_dim, const V* b,
                                 int64_t bs, V* residual, int64_t rs, V* gsin, int64_t sins,
                                 V* gcos, int64_t coss, uint8_t* stop)
{
    for (int64_t j = 0; j < cols; ++j) {
        for (int64_t i = 0; i < rows; ++i) residual[i * rs + j] = b[i * bs + j];
        for (int64_t i = 0; i < krylov_dim; ++i) gsin[i * sins + j] = gcos[i * coss + j] = 0;
        stop[j] = 0;
    }
}
/* reference/solver/common_gmres_kernels.cpp:27-107 (givens_rotation, calculate_sin_and_cos,
 * calculate_next_residual_norm) and :138-157 (hessenberg_qr) */
void FN(common_gmres_hessenberg_qr)(int64_t cols, V* gsin, int64_t sins, V* gcos, int64_t coss,
                                    V* residual_norm, V* rnc, int64_t rncs, V* hess, int64_t hs,
                                    int64_t iter, uint64_t* final_iter_nums, const uint8_t* stop)
{
    for (int64_t i = 0; i < cols; ++i)
        if (!st_has_stopped(stop[i])) final_iter_nums[i]++;
    for (int64_t i = 0; i < cols; ++i) {
        if (st_has_stopped(stop[i])) continue;
        for (int64_t j = 0; j < iter; ++j) {
            V temp = gcos[j * coss + i] * hess[j * hs + i] + gsin[j * sins + i] * hess[(j + 1) * hs + i];
            hess[(j + 1) * hs + i] =
                -gsin[j * sins + i] * hess[j * hs + i] + gcos[j * coss + i] * hess[(j + 1) * hs + i];
            hess[j * hs + i] = temp;
        }
        if (hess[iter * hs + i] == 0) {
            gcos[iter * coss + i] = 0;
            gsin[iter * sins + i] = 1;
        } else {
            V this_hess = hess[iter * hs + i];
            V next_hess = hess[(iter + 1) * hs + i];
            V scale = FABS(this_hess) + FABS(next_hess);
            V hyp = scale * SQRT(FABS(this_hess / scale) * FABS(this_hess / scale) +
                                 FABS(next_hess / scale) * FABS(next_hess / scale));
            gcos[iter * coss + i] = this_hess / hyp;
            gsin[iter * sins + i] = next_hess / hyp;
        }
        hess[iter * hs + i] =
            gcos[iter * coss + i] * hess[iter * hs + i] + gsin[iter * sins + i] * hess[(iter + 1) * hs + i];
        hess[(iter + 1) * hs + i] = 0;
    }
    for (int64_t i = 0; i < cols; ++i) {
        if (st_has_stopped(stop[i])) continue;
        rnc[(iter + 1) * rncs + i] = -gsin[iter * sins + i] * rnc[iter * rncs + i];
        rnc[iter * rncs + i] = gcos[iter * coss + i] * rnc[iter * rncs + i];
        residual_norm[i] = FABS(rnc[(iter + 1) * rncs + i]);
    }
}
/* reference/solver/common_gmres_kernels.cpp:162-188; H(i,j) at hess[j*hs + i*cols + k] */
void FN(common_gmres_solve_krylov)(int64_t cols, const V* rnc, int64_t rncs, const V* hess,
                                   int64_t hs, V* y, int64_t ys, const uint64_t* final_iter_nums,
                                   const uint8_t* stop)
{
    for (int64_t k = 0; k < cols; ++k) {
        if (st_is_finalized(stop[k])) continue;
        for (int64_t i = (int64_t)final_iter_nums[k] - 1; i >= 0; --i) {
            V temp = rnc[i * rncs + k];
            for (int64_t j = i + 1; j < (int64_t)final_iter_nums[k]; ++j)
                temp -= hess[j * hs + i * cols + k] * y[j * ys + k];
            y[i * ys + k] = temp / hess[i * hs + i * cols + k];
        }
    }
}
/* reference/solver/gmres_kernels.cpp:27-43 */
void FN(gmres_restart)(int64_t rows, int64_t cols, const V* residual, int64_t rs,
                       const V* residual_norm, V* rnc, V* krylov, int64_t ks,
                       uint64_t* final_iter_nums)
{
    for (int64_t j = 0; j < cols; ++j) {
        rnc[j] = residual_norm[j];
        for (int64_t i = 0; i < rows; ++i) krylov[i * ks + j] = residual[i * rs + j] / residual_norm[j];
        final_iter_nums[j] = 0;
    }
}
/* reference/solver/gmres_kernels.cpp:48-74 */
void FN(gmres_multi_axpy)(int64_t rows, int64_t cols, const V* krylov, int64_t ks, const V* y,
                          int64_t ys, V* out, int64_t os, const uint64_t* final_iter_nums,
                          uint8_t* stop)
{
    for (int64_t k = 0; k < cols; ++k) {
        if (st_is_finalized(stop[k])) continue;
        for (int64_t i = 0; i < rows; ++i) {
            out[i * os + k] = 0;
            for (int64_t j = 0; j < (int64_t)final_iter_nums[k]; ++j)
                out[i * os + k] += krylov[(i + j * rows) * ks + k] * y[j * ys + k];
        }
        if (st_has_stopped(stop[k])) stop[k] = st_finalize(stop[k]);
    }
}
/* reference/solver/gmres_kernels.cpp:79-98 */
void FN(gmres_multi_dot)(int64_t rows, int64_t cols, int64_t num_bases, const V* krylov,
                         int64_t ks, const V* next_krylov, int64_t ns, V* hcol, int64_t hs)
{
    for (int64_t i = 0; i < num_bases; ++i)
        for (int64_t k = 0; k < cols; ++k) {
            hcol[i * hs + k] = 0;
            for (int64_t j = 0; j < rows; ++j)
                hcol[i * hs + k] += krylov[(i * rows + j) * ks + k] * next_krylov[j * ns + k];
        }
}

/* reference/stop/residual_norm_kernels.cpp:27-55 (implicit: :68-92) */
void FN(residual_norm)(int64_t cols, const V* tau, const V* orig_tau, V goal, uint8_t id,
                       int32_t set_finalized, uint8_t* stop, uint8_t* device_storage,
                       int32_t* all_converged,
                       int32_t* one_changed)
{
    *all_converged = 1;
    *one_changed = 0;
    for (int64_t i = 0; i < cols; ++i)
        if (tau[i] <= goal * orig_tau[i]) {
            stop[i] = st_converge(stop[i], id, set_finalized);
            *one_changed = 1;
        }
    for (int64_t i = 0; i < cols; ++i)
        if (!st_has_stopped(stop[i])) {
            *all_converged = 0;
            break;
        }
}
void FN(implicit_residual_norm)(int64_t cols, const V* tau, const V* orig_tau, V goal, uint8_t id,
                                int32_t set_finalized, uint8_t* stop, uint8_t* device_storage,
                                int32_t* all_converged,
                                int32_t* one_changed)
{
    *all_converged = 1;
    *one_changed = 0;
    for (int64_t i = 0; i < cols; ++i)
        if (SQRT(FABS(tau[i])) <= goal * orig_tau[i]) {
            stop[i] = st_converge(stop[i], id, set_finalized);
            *one_changed = 1;
        }
    for (int64_t i = 0; i < cols; ++i)
        if (!st_has_stopped(stop[i])) {
            *all_converged = 0;
            break;
        }
}

/* reference/preconditioner/jacobi_kernels.cpp:577-592 */
void FN(jacobi_invert_diagonal)(int64_t n, const V* diag, V* inv_diag)
{
    for (int64_t i = 0; i < n; ++i) {
        V d = diag[i] == 0 ? (V)1 : diag[i];
        inv_diag[i] = (V)1 / d;
    }
}
/* reference/preconditioner/jacobi_kernels.cpp:541-555 */
void FN(jacobi_simple_scalar_apply)(int64_t rows, int64_t cols, const V* inv_diag, const V* b,
                                    int64_t bs, V* x, int64_t xs)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) x[i * xs + j] = b[i * bs + j] * inv_diag[i];
}
/* reference/preconditioner/jacobi_kernels.cpp:522-537 */
void FN(jacobi_scalar_apply)(int64_t rows, int64_t cols, const V* inv_diag, const V* alpha,
                             const V* b, int64_t bs, const V* beta, V* x, int64_t xs)
{
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j)
            x[i * xs + j] = beta[0] * x[i * xs + j] + alpha[0] * b[i * bs + j] * inv_diag[i];
}

#else /* ===================== (value, index)-typed kernels ===================== */

/* reference/matrix/csr_kernels.cpp:47-78 */
void FNI(csr_spmv)(int64_t num_rows, int64_t num_cols, int64_t nnz, const I* row_ptrs, const I* col_idxs, const V* values,
                   const V* b, int64_t bs, int64_t num_rhs, V* c, int64_t cs)
{
    for (int64_t row = 0; row < num_rows; ++row)
        for (int64_t j = 0; j < num_rhs; ++j) {
            V sum = 0;
            for (int64_t k = row_ptrs[row]; k < (int64_t)row_ptrs[row + 1]; ++k)
                sum += values[k] * b[(int64_t)col_idxs[k] * bs + j];
            c[row * cs + j] = sum;
        }
}
/* reference/matrix/csr_kernels.cpp:84-118 */
void FNI(csr_advanced_spmv)(int64_t num_rows, int64_t num_cols, int64_t nnz,
                            const I* row_ptrs, const I* col_idxs,
                            const V* values, const V* alpha, const V* b, int64_t bs,
                            int64_t num_rhs, const V* beta, V* c, int64_t cs)
{
    const V valpha = alpha[0], vbeta = beta[0];
    for (int64_t row = 0; row < num_rows; ++row)
        for (int64_t j = 0; j < num_rhs; ++j) {
            V sum = vbeta == 0 ? (V)0 : c[row * cs + j] * vbeta;
            for (int64_t k = row_ptrs[row]; k < (int64_t)row_ptrs[row + 1]; ++k)
                sum += valpha * values[k] * b[(int64_t)col_idxs[k] * bs + j];
            c[row * cs + j] = sum;
        }
}

/* reference/matrix/ell_kernels.cpp:29-72 */
void FNI(ell_spmv)(int64_t num_rows, int64_t num_cols, int64_t width, int64_t stride, const I* col_idxs,
                   const V* values, const V* b, int64_t bs, int64_t num_rhs, V* c, int64_t cs)
{
    for (int64_t j = 0; j < num_rhs; ++j)
        for (int64_t row = 0; row < num_rows; ++row) {
            V result = 0;
            for (int64_t i = 0; i < width; ++i) {
                I col = col_idxs[row + i * stride];
                if (col != (I)-1) result += values[row + i * stride] * b[(int64_t)col * bs + j];
            }
            c[row * cs + j] = result;
        }
}
/* reference/matrix/ell_kernels.cpp:77-120 */
void FNI(ell_advanced_spmv)(int64_t num_rows, int64_t num_cols, int64_t width, int64_t stride, const I* col_idxs,
                            const V* values, const V* alpha, const V* b, int64_t bs,
                            int64_t num_rhs, const V* beta, V* c, int64_t cs)
{
    const V a = alpha[0], bt = beta[0];
    for (int64_t j = 0; j < num_rhs; ++j)
        for (int64_t row = 0; row < num_rows; ++row) {
            V result = bt == 0 ? (V)0 : bt * c[row * cs + j];
            for (int64_t i = 0; i < width; ++i) {
                I col = col_idxs[row + i * stride];
                if (col != (I)-1)
                    result += a * values[row + i * stride] * b[(int64_t)col * bs + j];
            }
            c[row * cs + j] = result;
        }
}

/* reference/matrix/sellp_kernels.cpp:27-58 */
void FNI(sellp_spmv)(int64_t num_rows, int64_t num_cols, int64_t slice_size, const uint64_t* slice_sets,
                     const uint64_t* slice_lengths, const I* col_idxs, const V* values,
                     const V* b, int64_t bs, int64_t num_rhs, V* c, int64_t cs)
{
    int64_t slice_num = (num_rows + slice_size - 1) / slice_size;
    for (int64_t slice = 0; slice < slice_num; ++slice)
        for (int64_t row = 0; row < slice_size; ++row) {
            int64_t g = slice * slice_size + row;
            if (g >= num_rows) break;
            for (int64_t j = 0; j < num_rhs; ++j) c[g * cs + j] = 0;
            for (uint64_t i = 0; i < slice_lengths[slice]; ++i) {
                int64_t idx = (int64_t)(slice_sets[slice] + i) * slice_size + row;
                I col = col_idxs[idx];
                if (col != (I)-1)
                    for (int64_t j = 0; j < num_rhs; ++j)
                        c[g * cs + j] += values[idx] * b[(int64_t)col * bs + j];
            }
        }
}
/* reference/matrix/sellp_kernels.cpp:63-105 */
void FNI(sellp_advanced_spmv)(int64_t num_rows, int64_t num_cols, int64_t slice_size, const uint64_t* slice_sets,
                              const uint64_t* slice_lengths, const I* col_idxs, const V* values,
                              const V* alpha, const V* b, int64_t bs, int64_t num_rhs,
                              const V* beta, V* c, int64_t cs)
{
    const V va = alpha[0], vb = beta[0];
    int64_t slice_num = (num_rows + slice_size - 1) / slice_size;
    for (int64_t slice = 0; slice < slice_num; ++slice)
        for (int64_t row = 0; row < slice_size; ++row) {
            int64_t g = slice * slice_size + row;
            if (g >= num_rows) break;
            for (int64_t j = 0; j < num_rhs; ++j) {
                if (vb != 0)
                    c[g * cs + j] *= vb;
                else
                    c[g * cs + j] = 0;
            }
            for (uint64_t i = 0; i < slice_lengths[slice]; ++i) {
                int64_t idx = (int64_t)(slice_sets[slice] + i) * slice_size + row;
                I col = col_idxs[idx];
                if (col != (I)-1)
                    for (int64_t j = 0; j < num_rhs; ++j)
                        c[g * cs + j] += va * values[idx] * b[(int64_t)col * bs + j];
            }
        }
}

/* reference/matrix/coo_kernels.cpp:59-74 (spmv2), :81-97 (advanced_spmv2),
 * :33-40 (spmv = fill 0 + spmv2), :46-55 (advanced_spmv = scale(beta) + advanced_spmv2) */
void FNI(coo_spmv2)(int64_t num_rows, int64_t num_cols, int64_t nnz, const I* row_idxs, const I* col_idxs, const V* values,
                    const V* b, int64_t bs, int64_t num_rhs, V* c, int64_t cs)
{
    for (int64_t i = 0; i < nnz; ++i)
        for (int64_t j = 0; j < num_rhs; ++j)
            c[(int64_t)row_idxs[i] * cs + j] += values[i] * b[(int64_t)col_idxs[i] * bs + j];
}
void FNI(coo_advanced_spmv2)(int64_t num_rows, int64_t num_cols, int64_t nnz,
                             const I* row_idxs, const I* col_idxs, const V* values,
                             const V* alpha, const V* b, int64_t bs, int64_t num_rhs, V* c,
                             int64_t cs)
{
    const V a = alpha[0];
    for (int64_t i = 0; i < nnz; ++i)
        for (int64_t j = 0; j < num_rhs; ++j)
            c[(int64_t)row_idxs[i] * cs + j] += a * values[i] * b[(int64_t)col_idxs[i] * bs + j];
}
void FNI(coo_spmv)(int64_t num_rows, int64_t num_cols, int64_t nnz, const I* row_idxs, const I* col_idxs,
                   const V* values, const V* b, int64_t bs, int64_t num_rhs, V* c, int64_t cs)
{
    for (int64_t i = 0; i < num_rows; ++i)
        for (int64_t j = 0; j < num_rhs; ++j) c[i * cs + j] = 0;
    FNI(coo_spmv2)(num_rows, num_cols, nnz, row_idxs, col_idxs, values, b, bs, num_rhs, c, cs);
}
void FNI(coo_advanced_spmv)(int64_t num_rows, int64_t num_cols, int64_t nnz, const I* row_idxs, const I* col_idxs,
                            const V* values, const V* alpha, const V* b, int64_t bs,
                            int64_t num_rhs, const V* beta, V* c, int64_t cs)
{
    FN(dense_scale)(num_rows, num_rhs, beta, 1, c, cs);
    FNI(coo_advanced_spmv2)(num_rows, num_cols, nnz, row_idxs, col_idxs, values, alpha, b, bs,
                            num_rhs, c, cs);
}

/* reference/matrix/csr_kernels.cpp `extract_diagonal`: first stored entry with col == row */
void FNI(csr_extract_diagonal)(int64_t n, const I* rp, const I* ci, const V* va, V* diag)
{
    for (int64_t row = 0; row < n; ++row) {
        diag[row] = 0;
        for (int64_t k = rp[row]; k < (int64_t)rp[row + 1]; ++k)
            if ((int64_t)ci[k] == row) {
                diag[row] = va[k];
                break;
            }
    }
}

/* reference/matrix/csr_kernels.cpp:573-598 convert_to_ell: every (row, i < width) slot is
 * reset to (0, -1) and the row's entries copied in order; rows >= num_rows of a padded
 * stride are not touched. */
void FNI(csr_convert_to_ell)(int64_t num_rows, const I* rp, const I* ci, const V* va,
                             int64_t width, int64_t stride, I* ecols, V* evals)
{
    for (int64_t row = 0; row < num_rows; ++row) {
        for (int64_t i = 0; i < width; ++i) {
            evals[row + i * stride] = 0;
            ecols[row + i * stride] = (I)-1;
        }
        for (int64_t i = 0; i < (int64_t)rp[row + 1] - (int64_t)rp[row]; ++i) {
            evals[row + i * stride] = va[rp[row] + i];
            ecols[row + i * stride] = ci[rp[row] + i];
        }
    }
}

/* reference/matrix/csr_kernels.cpp:528-567 convert_to_sellp */
void FNI(csr_convert_to_sellp)(int64_t num_rows, int64_t slice_size, const uint64_t* slice_sets,
                               const uint64_t* slice_lengths, const I* rp, const I* ci,
                               const V* va, I* scols, V* svals)
{
    const int64_t slice_num = (num_rows + slice_size - 1) / slice_size;
    for (int64_t slice = 0; slice < slice_num; ++slice) {
        for (int64_t row = 0; row < slice_size; ++row) {
            const int64_t global_row = slice * slice_size + row;
            if (global_row >= num_rows) break;
            int64_t ind = (int64_t)slice_sets[slice] * slice_size + row;
            for (int64_t k = rp[global_row]; k < (int64_t)rp[global_row + 1]; ++k) {
                svals[ind] = va[k];
                scols[ind] = ci[k];
                ind += slice_size;
            }
            for (int64_t i = ind;
                 i < (int64_t)(slice_sets[slice] + slice_lengths[slice]) * slice_size + row;
                 i += slice_size) {
                scols[i] = (I)-1;
                svals[i] = 0;
            }
        }
    }
}

/* reference/matrix/csr_kernels.cpp:910-953 convert_to_hybrid: the whole ELL part
 * (ell_stride rows x ell_lim) is reset, then each row fills ELL first and spills into COO */
void FNI(csr_convert_to_hybrid)(int64_t num_rows, const I* rp, const I* ci, const V* va,
                                int64_t ell_lim, int64_t ell_stride, I* ecols, V* evals,
                                const int64_t* coo_row_ptrs, I* crows, I* ccols, V* cvals)
{
    (void)coo_row_ptrs; /* the reference kernel ignores them as well (sequential counter) */
    for (int64_t i = 0; i < ell_lim; ++i)
        for (int64_t j = 0; j < ell_stride; ++j) {
            evals[j + i * ell_stride] = 0;
            ecols[j + i * ell_stride] = (I)-1;
        }
    int64_t csr_idx = 0, coo_idx = 0;
    for (int64_t row = 0; row < num_rows; ++row) {
        int64_t ell_idx = 0;
        while (csr_idx < (int64_t)rp[row + 1]) {
            if (ell_idx < ell_lim) {
                evals[row + ell_idx * ell_stride] = va[csr_idx];
                ecols[row + ell_idx * ell_stride] = ci[csr_idx];
                ++ell_idx;
            } else {
                cvals[coo_idx] = va[csr_idx];
                ccols[coo_idx] = ci[csr_idx];
                crows[coo_idx] = (I)row;
                ++coo_idx;
            }
            ++csr_idx;
        }
    }
}

/* reference/matrix/csr_kernels.cpp:1272-1290 sort_by_column_index: std::sort of the
 * (col, value) pairs of each row by column.  Restated as a stable insertion sort; identical
 * for rows with distinct columns (all that an unstable sort defines). */
void FNI(csr_sort_by_column_index)(int64_t num_rows, const I* rp, I* ci, V* va)
{
    for (int64_t row = 0; row < num_rows; ++row) {
        const int64_t s = rp[row], e = rp[row + 1];
        for (int64_t i = s + 1; i < e; ++i) {
            const I c = ci[i];
            const V v = va[i];
            int64_t j = i - 1;
            while (j >= s && ci[j] > c) {
                ci[j + 1] = ci[j];
                va[j + 1] = va[j];
                --j;
            }
            ci[j + 1] = c;
            va[j + 1] = v;
        }
    }
}

/* reference/preconditioner/jacobi_kernels.cpp:419-447 (apply_block), :460-520 (apply /
 * simple_apply); storage scheme include/ginkgo/core/preconditioner/jacobi.hpp:37-141 */
void FNI(jacobi_apply)(int64_t num_blocks, int32_t max_block_size, int64_t block_offset, int64_t group_offset,
                       int32_t group_power, const I* block_ptrs, const V* blocks, const V* alpha_p,
                       const V* b, int64_t bs, int64_t num_rhs, const V* beta_p, V* x, int64_t xs)
{
    const V alpha = alpha_p ? alpha_p[0] : (V)1;
    const V beta = beta_p ? beta_p[0] : (V)0;
    const int64_t stride = block_offset << group_power;
    for (int64_t k = 0; k < num_blocks; ++k) {
        const V* blk = blocks + group_offset * (k >> group_power) +
                       block_offset * (k & (((int64_t)1 << group_power) - 1));
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
            for (int64_t row = 0; row < n; ++row)
                for (int64_t col = 0; col < num_rhs; ++col)
                    xb[row * xs + col] += alpha * blk[row + inner * stride] * bb[inner * bs + col];
    }
}


/* reference/preconditioner/jacobi_kernels.cpp:113-147 extract_block, :150-205 choose_pivot /
 * swap_rows / apply_gauss_jordan_transform, :262-278 invert_block, :243-258
 * permute_and_transpose_block, :320-410 generate (full-precision branch, no conditioning) */
void FNI(jacobi_generate)(int64_t num_rows, const I* rp, const I* ci, const V* va, int64_t num_blocks,
                          int32_t max_block_size, int64_t block_offset, int64_t group_offset,
                          int32_t group_power, const I* block_ptrs, V* blocks)
{
    (void)num_rows;
    (void)max_block_size;
    const int64_t stride = block_offset << group_power;
    V blk[32 * 32];
    int perm[32];
    for (int64_t g = 0; g < num_blocks; ++g) {
        const int64_t start = block_ptrs[g];
        const int bs = (int)((int64_t)block_ptrs[g + 1] - start);
        for (int i = 0; i < bs; ++i)
            for (int j = 0; j < bs; ++j) blk[i * bs + j] = 0;
        for (int i = 0; i < bs; ++i) perm[i] = i;
        for (int row = 0; row < bs; ++row)
            for (int64_t p = rp[start + row]; p < (int64_t)rp[start + row + 1]; ++p) {
                const int64_t col = (int64_t)ci[p] - start;
                if (0 <= col && col < bs) blk[row * bs + col] = va[p];
            }
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
            if (d == 0) break;
            for (int i = 0; i < bs; ++i) blk[i * bs + k] /= -d;
            blk[k * bs + k] = 0;
            for (int i = 0; i < bs; ++i)
                for (int j = 0; j < bs; ++j) blk[i * bs + j] += blk[i * bs + k] * blk[k * bs + j];
            for (int j = 0; j < bs; ++j) blk[k * bs + j] /= d;
            blk[k * bs + k] = (V)1 / d;
        }
        V* dst = blocks + group_offset * (g >> group_power) +
                 block_offset * (g & (((int64_t)1 << group_power) - 1));
        for (int i = 0; i < bs; ++i)
            for (int j = 0; j < bs; ++j) dst[i + perm[j] * stride] = blk[i * bs + j];
    }
}

/* reference/preconditioner/jacobi_kernels.cpp:493-520 (simple_apply == alpha 1, beta 0) */
void FNI(jacobi_simple_apply)(int64_t num_blocks, int32_t max_block_size, int64_t block_offset,
                              int64_t group_offset, int32_t group_power, const I* block_ptrs,
                              const V* blocks, const V* b, int64_t bs, int64_t num_rhs, V* x,
                              int64_t xs)
{
    FNI(jacobi_apply)(num_blocks, max_block_size, block_offset, group_offset, group_power,
                      block_ptrs, blocks, NULL, b, bs, num_rhs, NULL, x, xs);
}

/* reference/matrix/csr_kernels.cpp:694-719 transpose_and_transform with the identity: count the
 * columns, prefix sum, then walk the rows in order (convert_csr_to_csc) -- inside a row of the
 * transpose the entries are ordered by (original row, position) */
void FNI(csr_transpose)(int64_t num_rows, int64_t num_cols, int64_t nnz, const I* row_ptrs, const I* col_idxs,
                        const V* values, I* t_row_ptrs, I* t_col_idxs, V* t_values)
{
    (void)nnz;
    for (int64_t c = 0; c <= num_cols; ++c) t_row_ptrs[c] = 0;
    for (int64_t k = 0; k < (int64_t)row_ptrs[num_rows]; ++k) t_row_ptrs[col_idxs[k] + 1]++;
    for (int64_t c = 0; c < num_cols; ++c) t_row_ptrs[c + 1] += t_row_ptrs[c];
    I* next = (I*)malloc(sizeof(I) * (size_t)(num_cols > 0 ? num_cols : 1));
    for (int64_t c = 0; c < num_cols; ++c) next[c] = t_row_ptrs[c];
    for (int64_t row = 0; row < num_rows; ++row)
        for (int64_t k = row_ptrs[row]; k < (int64_t)row_ptrs[row + 1]; ++k) {
            const I dest = next[col_idxs[k]]++;
            t_col_idxs[dest] = (I)row;
            t_values[dest] = values[k];
        }
    free(next);
}

/* reference/preconditioner/jacobi_kernels.cpp:208-218 transpose_block, :597-627 transpose_jacobi
 * (full-precision storage): out(j, i) = in(i, j) inside every block, same storage scheme */
void FNI(jacobi_transpose)(int64_t num_blocks, int32_t max_block_size, int64_t block_offset,
                           int64_t group_offset, int32_t group_power, const I* block_ptrs, const V* blocks,
                           V* out_blocks)
{
    (void)max_block_size;
    const int64_t stride = block_offset << group_power;
    for (int64_t k = 0; k < num_blocks; ++k) {
        const int64_t ofs = group_offset * (k >> group_power) +
                            block_offset * (k & (((int64_t)1 << group_power) - 1));
        const int64_t n = (int64_t)block_ptrs[k + 1] - block_ptrs[k];
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) out_blocks[ofs + i * stride + j] = blocks[ofs + i + j * stride];
    }
}

/* adaptive-precision block-Jacobi (storage_optimization): generate / apply / transpose */
#include "oracle_jacobi_adaptive.h"

#endif
