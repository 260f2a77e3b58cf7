// Drop-in Krylov step kernels: CG, BiCGStab (element-wise, masked by the
// per-column stopping_status) -- replaces gko::kernels::cuda::{cg,bicgstab}::*
// (reference common/unified/solver/{cg,bicgstab}_kernels.cpp); arithmetic
// contract reference/solver/cg_kernels.cpp:24-102 and
// reference/solver/bicgstab_kernels.cpp:25-178.  Scalars (rho, beta, ...) are
// read from device memory by every thread, no host round trip.
#include "elementwise.cuh"

namespace b200 {
namespace steps {

template <typename V>
b200_status cg_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs, V* r,
                          int64_t rs, V* z, int64_t zs, V* p, int64_t ps, V* q, int64_t qs,
                          V* prev_rho, V* rho, uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    // the scalar rows are initialised by the first `cols` threads of an extra row
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            rho[j] = V(0);
            prev_rho[j] = V(1);
            stop[j] = 0;
        } else {
            r[i * rs + j] = b[i * bs + j];
            z[i * zs + j] = V(0);
            p[i * ps + j] = V(0);
            q[i * qs + j] = V(0);
        }
    });
}

template <typename V>
b200_status cg_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, V* p, int64_t ps, const V* z,
                      int64_t zs, const V* rho, const V* prev_rho, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V pr = prev_rho[j];
        if (pr == V(0)) {
            p[i * ps + j] = z[i * zs + j];
        } else {
            const V tmp = rho[j] / pr;
            p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
        }
    });
}

template <typename V>
b200_status cg_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* r,
                      int64_t rs, const V* p, int64_t ps, const V* q, int64_t qs, const V* beta,
                      const V* rho, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = beta[j];
        if (bt != V(0)) {
            const V tmp = rho[j] / bt;
            x[i * xs + j] += tmp * p[i * ps + j];
            r[i * rs + j] -= tmp * q[i * qs + j];
        }
    });
}

// ---- FCG (reference/solver/fcg_kernels.cpp:22-100) and CGS (reference/solver/
// cgs_kernels.cpp:22-140): siblings of CG with the same element-wise + device-scalar pattern
template <typename V>
b200_status fcg_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs, V* r,
                           int64_t rs, V* z, int64_t zs, V* p, int64_t ps, V* q, int64_t qs, V* t,
                           int64_t ts, V* prev_rho, V* rho, V* rho_t, uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            rho[j] = V(0);
            prev_rho[j] = V(1);
            rho_t[j] = V(1);
            stop[j] = 0;
        } else {
            const V v = b[i * bs + j];
            t[i * ts + j] = v;
            r[i * rs + j] = v;
            z[i * zs + j] = V(0);
            p[i * ps + j] = V(0);
            q[i * qs + j] = V(0);
        }
    });
}

template <typename V>
b200_status fcg_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, V* p, int64_t ps, const V* z,
                       int64_t zs, const V* rho_t, const V* prev_rho, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V pr = prev_rho[j];
        if (pr == V(0)) {
            p[i * ps + j] = z[i * zs + j];
        } else {
            const V tmp = rho_t[j] / pr;
            p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
        }
    });
}

template <typename V>
b200_status fcg_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* r,
                       int64_t rs, V* t, int64_t ts, const V* p, int64_t ps, const V* q,
                       int64_t qs, const V* beta, const V* rho, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = beta[j];
        if (bt != V(0)) {
            const V tmp = rho[j] / bt;
            const V prev_r = r[i * rs + j];
            x[i * xs + j] += tmp * p[i * ps + j];
            const V nr = prev_r - tmp * q[i * qs + j];
            r[i * rs + j] = nr;
            t[i * ts + j] = nr - prev_r;
        }
    });
}

template <typename V>
b200_status cgs_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs, V* r,
                           int64_t rs, V* r_tld, int64_t rts, V* p, int64_t ps, V* q, int64_t qs,
                           V* u, int64_t us, V* u_hat, int64_t uhs, V* v_hat, int64_t vhs, V* t,
                           int64_t ts, V* alpha, V* beta, V* gamma, V* prev_rho, V* rho,
                           uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            rho[j] = V(0);
            prev_rho[j] = V(1);
            alpha[j] = V(1);
            beta[j] = V(1);
            gamma[j] = V(1);
            stop[j] = 0;
        } else {
            const V v = b[i * bs + j];
            r[i * rs + j] = v;
            r_tld[i * rts + j] = v;
            u[i * us + j] = V(0);
            u_hat[i * uhs + j] = V(0);
            p[i * ps + j] = V(0);
            q[i * qs + j] = V(0);
            v_hat[i * vhs + j] = V(0);
            t[i * ts + j] = V(0);
        }
    });
}

// step_1 and step_2 also WRITE a 1 x cols scalar (beta resp. alpha) that every thread of the
// column needs: each thread recomputes it from the inputs, and the store is done by a
// separate tiny launch BEFORE the vector update, exactly the reference's two loops.
template <typename V>
b200_status cgs_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, const V* r, int64_t rs, V* u,
                       int64_t us, V* p, int64_t ps, const V* q, int64_t qs, V* beta,
                       const V* rho, const V* prev_rho, const uint8_t* stop)
{
    b200_status st = launch_ew(ctx, 1, cols, [=] __device__(int64_t, int64_t j) {
        if (has_stopped(stop[j])) return;
        if (prev_rho[j] != V(0)) beta[j] = rho[j] / prev_rho[j];
    });
    if (st != B200_OK) return st;
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = beta[j];
        const V qv = q[i * qs + j];
        const V uv = r[i * rs + j] + bt * qv;
        u[i * us + j] = uv;
        p[i * ps + j] = uv + bt * (qv + bt * p[i * ps + j]);
    });
}

template <typename V>
b200_status cgs_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, const V* u, int64_t us,
                       const V* v_hat, int64_t vhs, V* q, int64_t qs, V* t, int64_t ts, V* alpha,
                       const V* rho, const V* gamma, const uint8_t* stop)
{
    b200_status st = launch_ew(ctx, 1, cols, [=] __device__(int64_t, int64_t j) {
        if (has_stopped(stop[j])) return;
        if (gamma[j] != V(0)) alpha[j] = rho[j] / gamma[j];
    });
    if (st != B200_OK) return st;
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V uv = u[i * us + j];
        const V qv = uv - alpha[j] * v_hat[i * vhs + j];
        q[i * qs + j] = qv;
        t[i * ts + j] = uv + qv;
    });
}

template <typename V>
b200_status cgs_step_3(b200_ctx* ctx, int64_t rows, int64_t cols, const V* t, int64_t ts,
                       const V* u_hat, int64_t uhs, V* r, int64_t rs, V* x, int64_t xs,
                       const V* alpha, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V a = alpha[j];
        x[i * xs + j] += a * u_hat[i * uhs + j];
        r[i * rs + j] -= a * t[i * ts + j];
    });
}

// ---- IR / Chebyshev (reference/solver/ir_kernels.cpp:17-24, chebyshev_kernels.cpp:15-66).
// The Chebyshev coefficients arrive by value and the arithmetic is done in double whatever V
// is (solver::detail::coeff_type, include/ginkgo/core/solver/chebyshev.hpp:29-31).
template <typename V>
b200_status chebyshev_init_update(b200_ctx* ctx, int64_t rows, int64_t cols, double alpha,
                                  const V* inner_sol, int64_t is, V* update_sol, int64_t us,
                                  V* output, int64_t os)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        const double inner_val = (double)inner_sol[i * is + j];
        update_sol[i * us + j] = (V)inner_val;
        const double prod = alpha * inner_val;
        output[i * os + j] = (V)((double)output[i * os + j] + prod);
    });
}

template <typename V>
b200_status chebyshev_update(b200_ctx* ctx, int64_t rows, int64_t cols, double alpha, double beta,
                             V* inner_sol, int64_t is, V* update_sol, int64_t us, V* output,
                             int64_t os)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        const double bu = beta * (double)update_sol[i * us + j];
        const double val = (double)inner_sol[i * is + j] + bu;
        inner_sol[i * is + j] = (V)val;
        update_sol[i * us + j] = (V)val;
        const double prod = alpha * val;
        output[i * os + j] = (V)((double)output[i * os + j] + prod);
    });
}

// ---- PipeCG (reference/solver/pipe_cg_kernels.cpp:24-160)
template <typename V>
b200_status pipe_cg_initialize_1(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs,
                                 V* r, int64_t rs, V* prev_rho, uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            prev_rho[j] = V(1);
            stop[j] = 0;
        } else {
            r[i * rs + j] = b[i * bs + j];
        }
    });
}

template <typename V>
b200_status pipe_cg_initialize_2(b200_ctx* ctx, int64_t rows, int64_t cols, V* p, int64_t ps, V* q,
                                 int64_t qs, V* f, int64_t fs, V* g, int64_t gs, V* beta, const V* z,
                                 int64_t zs, const V* w, int64_t ws, const V* m, int64_t ms,
                                 const V* n, int64_t ns, const V* delta)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            beta[j] = delta[j];
        } else {
            p[i * ps + j] = z[i * zs + j];
            q[i * qs + j] = w[i * ws + j];
            f[i * fs + j] = m[i * ms + j];
            g[i * gs + j] = n[i * ns + j];
        }
    });
}

template <typename V>
b200_status pipe_cg_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* r,
                           int64_t rs, V* z1, int64_t z1s, V* z2, int64_t z2s, V* w, int64_t ws,
                           const V* p, int64_t ps, const V* q, int64_t qs, const V* f, int64_t fs,
                           const V* g, int64_t gs, const V* rho, const V* beta, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = beta[j];
        if (bt != V(0)) {
            const V tmp = rho[j] / bt;
            x[i * xs + j] += tmp * p[i * ps + j];
            r[i * rs + j] -= tmp * q[i * qs + j];
            const V z = z1[i * z1s + j] - tmp * f[i * fs + j];
            z1[i * z1s + j] = z;
            z2[i * z2s + j] = z;
            w[i * ws + j] -= tmp * g[i * gs + j];
        }
    });
}

// step_2 rewrites beta (1 x cols) from its own old value, so the scalar update runs first in a
// tiny launch that also leaves tmp = rho / prev_rho (or the "restart" flag) for the vector pass
template <typename V>
b200_status pipe_cg_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, V* beta, V* p, int64_t ps, V* q,
                           int64_t qs, V* f, int64_t fs, V* g, int64_t gs, const V* z, int64_t zs,
                           const V* w, int64_t ws, const V* m, int64_t ms, const V* n, int64_t ns,
                           const V* prev_rho, const V* rho, const V* delta, const uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    b200_status st = launch_ew(ctx, 1, cols, [=] __device__(int64_t, int64_t j) {
        if (has_stopped(stop[j])) return;
        if (prev_rho[j] != V(0)) {
            const V tmp = rho[j] / prev_rho[j];
            const V abs_tmp = tmp < V(0) ? -tmp : tmp;
            const V sq = abs_tmp * abs_tmp;
            V bt = delta[j] - sq * beta[j];
            if (bt == V(0)) bt = delta[j];
            beta[j] = bt;
        } else {
            beta[j] = delta[j];
        }
    });
    if (st != B200_OK) return st;
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V pr = prev_rho[j];
        if (pr != V(0)) {
            const V tmp = rho[j] / pr;
            p[i * ps + j] = z[i * zs + j] + tmp * p[i * ps + j];
            q[i * qs + j] = w[i * ws + j] + tmp * q[i * qs + j];
            f[i * fs + j] = m[i * ms + j] + tmp * f[i * fs + j];
            g[i * gs + j] = n[i * ns + j] + tmp * g[i * gs + j];
        } else {
            p[i * ps + j] = z[i * zs + j];
            q[i * qs + j] = w[i * ws + j];
            f[i * fs + j] = m[i * ms + j];
            g[i * gs + j] = n[i * ns + j];
        }
    });
}

// ---- GCR (reference/solver/gcr_kernels.cpp:26-84); the orthogonalisation itself is built from
// the Dense kernels (dot, squared_norm2, inv_scale, sub_scaled) by the host loop, as in the reference
template <typename V>
b200_status gcr_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs,
                           V* residual, int64_t rs, uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows)
            stop[j] = 0;
        else
            residual[i * rs + j] = b[i * bs + j];
    });
}

template <typename V>
b200_status gcr_restart(b200_ctx* ctx, int64_t rows, int64_t cols, const V* residual, int64_t rs,
                        const V* a_residual, int64_t ars, V* p_bases, int64_t ps, V* ap_bases,
                        int64_t aps, uint64_t* final_iter_nums)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            final_iter_nums[j] = 0;
        } else {
            p_bases[i * ps + j] = residual[i * rs + j];
            ap_bases[i * aps + j] = a_residual[i * ars + j];
        }
    });
}

template <typename V>
b200_status gcr_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* residual,
                       int64_t rs, const V* p, int64_t ps, const V* ap, int64_t aps,
                       const V* ap_norm, const V* rap, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V an = ap_norm[j];
        if (an != V(0)) {
            const V tmp = rap[j] / an;
            x[i * xs + j] += tmp * p[i * ps + j];
            residual[i * rs + j] -= tmp * ap[i * aps + j];
        }
    });
}

// ---- MINRES (reference/solver/minres_kernels.cpp:26-150).  safe_divide(a, b) = b == 0 ? 0 : a / b
// (include/ginkgo/core/base/math.hpp); sqrt and division are IEEE on the device, products and
// sums are rounded separately (-fmad=false), so the scalar recurrences match the reference.
template <typename V>
__device__ __forceinline__ V safe_divide(V a, V b)
{
    return b == V(0) ? V(0) : a / b;
}
template <typename V>
__device__ __forceinline__ V dev_abs(V a)
{
    return a < V(0) ? -a : a;
}

template <typename V>
b200_status minres_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* r, int64_t rs, V* z,
                              int64_t zs, V* p, int64_t ps, V* p_prev, int64_t pps, V* q, int64_t qs,
                              V* q_prev, int64_t qps, V* q_tilde, int64_t qts, V* beta, V* gamma,
                              V* delta, V* cos_prev, V* cos_, V* sin_prev, V* sin_, V* eta_next, V* eta,
                              uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    // scalars first (beta <- sqrt(beta)), then the vectors read the new beta
    b200_status st = launch_ew(ctx, 1, cols, [=] __device__(int64_t, int64_t j) {
        delta[j] = V(0);
        gamma[j] = V(0);
        cos_prev[j] = V(0);
        sin_prev[j] = V(0);
        sin_[j] = V(0);
        cos_[j] = V(1);
        const V sb = sqrt(beta[j]);
        beta[j] = sb;
        eta[j] = sb;
        eta_next[j] = sb;
        stop[j] = 0;
    });
    if (st != B200_OK) return st;
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        const V bt = beta[j];
        q[i * qs + j] = safe_divide(r[i * rs + j], bt);
        z[i * zs + j] = safe_divide(z[i * zs + j], bt);
        p[i * ps + j] = V(0);
        p_prev[i * pps + j] = V(0);
        q_prev[i * qps + j] = V(0);
        q_tilde[i * qts + j] = V(0);
    });
}

template <typename V>
b200_status minres_step_1(b200_ctx* ctx, int64_t cols, V* alpha, V* beta, V* gamma, V* delta,
                          V* cos_prev, V* cos_, V* sin_prev, V* sin_, V* eta, V* eta_next, V* tau,
                          const uint8_t* stop)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return launch_ew(ctx, 1, cols, [=] __device__(int64_t, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = sqrt(beta[j]);
        beta[j] = bt;
        const V tmp_d = gamma[j];
        const V tmp_a = alpha[j];
        delta[j] = sin_prev[j] * tmp_d;
        const V c_old = cos_[j], s_old = sin_[j], cp_old = cos_prev[j];
        const V g1 = cp_old * c_old;
        const V g2 = g1 * tmp_d;
        const V g3 = s_old * tmp_a;
        gamma[j] = g2 + g3;
        const V a1 = (-s_old) * cp_old;
        const V a2 = a1 * tmp_d;
        const V a3 = c_old * tmp_a;
        V al = a2 + a3;
        // swap(cos, cos_prev), swap(sin, sin_prev): the old pair moves to *_prev
        cos_prev[j] = c_old;
        sin_prev[j] = s_old;
        V c, sn;
        if (al == V(0)) {
            c = V(0);
            sn = V(1);
        } else {
            const V scale = dev_abs(al) + dev_abs(bt);
            const V ra = dev_abs(al / scale), rb = dev_abs(bt / scale);
            const V h1 = ra * ra;
            const V h2 = rb * rb;
            const V hyp = scale * sqrt(h1 + h2);
            c = al / hyp;
            sn = bt / hyp;
        }
        const V n1 = c * al;
        const V n2 = sn * bt;
        al = n1 + n2;
        alpha[j] = al;
        cos_[j] = c;
        sin_[j] = sn;
        const V t1 = sn * sn;
        tau[j] = t1 * tau[j];
        const V e = eta_next[j];
        eta[j] = e;
        eta_next[j] = (-sn) * e;
    });
}

template <typename V>
b200_status minres_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* p, int64_t ps,
                          const V* p_prev, int64_t pps, V* z, int64_t zs, const V* z_tilde, int64_t zts,
                          V* q, int64_t qs, V* q_prev, int64_t qps, V* v, int64_t vs, const V* alpha,
                          const V* beta, const V* gamma, const V* delta, const V* cos_, const V* eta,
                          const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V gp = gamma[j] * p_prev[i * pps + j];
        const V dp = delta[j] * p[i * ps + j];
        const V num = (z[i * zs + j] - gp) - dp;
        const V pn = safe_divide(num, alpha[j]);
        p[i * ps + j] = pn;
        const V ce = cos_[j] * eta[j];
        const V cep = ce * pn;
        x[i * xs + j] = x[i * xs + j] + cep;
        const V vv = v[i * vs + j];
        q_prev[i * qps + j] = vv;
        const V tmp = q[i * qs + j];
        const V bt = beta[j];
        q[i * qs + j] = safe_divide(vv, bt);
        v[i * vs + j] = tmp * bt;
        z[i * zs + j] = safe_divide(z_tilde[i * zts + j], bt);
    });
}

template <typename V>
b200_status bicgstab_initialize(b200_ctx* ctx, int64_t rows, int64_t cols, const V* b, int64_t bs,
                                V* r, int64_t rs, V* rr, int64_t rrs, V* y, int64_t ys, V* s,
                                int64_t ss, V* t, int64_t ts, V* z, int64_t zs, V* v, int64_t vs,
                                V* p, int64_t ps, V* prev_rho, V* rho, V* alpha, V* beta,
                                V* gamma, V* omega, uint8_t* stop)
{
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (i == rows) {
            rho[j] = V(1);
            prev_rho[j] = V(1);
            alpha[j] = V(1);
            beta[j] = V(1);
            gamma[j] = V(1);
            omega[j] = V(1);
            stop[j] = 0;
        } else {
            r[i * rs + j] = b[i * bs + j];
            rr[i * rrs + j] = V(0);
            z[i * zs + j] = V(0);
            v[i * vs + j] = V(0);
            s[i * ss + j] = V(0);
            t[i * ts + j] = V(0);
            y[i * ys + j] = V(0);
            p[i * ps + j] = V(0);
        }
    });
}

template <typename V>
b200_status bicgstab_step_1(b200_ctx* ctx, int64_t rows, int64_t cols, const V* r, int64_t rs,
                            V* p, int64_t ps, const V* v, int64_t vs, const V* rho,
                            const V* prev_rho, const V* alpha, const V* omega,
                            const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V pr = prev_rho[j], om = omega[j];
        if (pr * om != V(0)) {
            const V tmp = rho[j] / pr * alpha[j] / om;
            p[i * ps + j] = r[i * rs + j] + tmp * (p[i * ps + j] - om * v[i * vs + j]);
        } else {
            p[i * ps + j] = r[i * rs + j];
        }
    });
}

// step_2 also WRITES alpha (1 x cols).  Every thread of a column computes the
// same value; row 0's thread stores it after all reads of the old alpha are
// irrelevant (alpha is not an input of this kernel).
template <typename V>
b200_status bicgstab_step_2(b200_ctx* ctx, int64_t rows, int64_t cols, const V* r, int64_t rs,
                            V* s, int64_t ss, const V* v, int64_t vs, const V* rho, V* alpha,
                            const V* beta, const uint8_t* stop)
{
    return launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = beta[j];
        V a = V(0);
        if (bt != V(0)) {
            a = rho[j] / bt;
            s[i * ss + j] = r[i * rs + j] - a * v[i * vs + j];
        } else {
            s[i * ss + j] = r[i * rs + j];
        }
        if (i == 0) alpha[j] = a;
    });
}

// step_3 also WRITES omega, which is not read by this kernel's other threads
// (they recompute it from gamma/beta).
template <typename V>
b200_status bicgstab_step_3(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs, V* r,
                            int64_t rs, const V* s, int64_t ss, const V* t, int64_t ts,
                            const V* y, int64_t ys, const V* z, int64_t zs, const V* alpha,
                            const V* beta, const V* gamma, V* omega, const uint8_t* stop)
{
    return launch_ew(ctx, rows + 1, cols, [=] __device__(int64_t i, int64_t j) {
        if (has_stopped(stop[j])) return;
        const V bt = beta[j];
        const V om = bt != V(0) ? gamma[j] / bt : V(0);
        if (i == rows) {
            omega[j] = om;
            return;
        }
        x[i * xs + j] += alpha[j] * y[i * ys + j] + om * z[i * zs + j];
        r[i * rs + j] = s[i * ss + j] - om * t[i * ts + j];
    });
}

template <typename V>
__global__ void finalize_status_kernel(int64_t cols, uint8_t* stop)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < cols && has_stopped(stop[j])) stop[j] |= kFinalizedMask;
}

template <typename V>
b200_status bicgstab_finalize(b200_ctx* ctx, int64_t rows, int64_t cols, V* x, int64_t xs,
                              const V* y, int64_t ys, const V* alpha, uint8_t* stop)
{
    b200_status st = launch_ew(ctx, rows, cols, [=] __device__(int64_t i, int64_t j) {
        const uint8_t sj = stop[j];
        if (has_stopped(sj) && !is_finalized(sj)) x[i * xs + j] += alpha[j] * y[i * ys + j];
    });
    if (st != B200_OK || cols == 0) return st;
    // the status flip must come after every thread has read the old status
    if (rows > 0) {
        finalize_status_kernel<V><<<(unsigned)ceildiv(cols, 256), 256, 0, ctx->stream>>>(cols, stop);
        B200_LAUNCH_CHECK(ctx);
    }
    return B200_OK;
}

}  // namespace steps
}  // namespace b200

extern "C" {

/* ir::initialize (reference/solver/ir_kernels.cpp:17-24): reset every stopping status */
b200_status b200_ir_initialize(b200_ctx* ctx, int64_t cols, uint8_t* stop_status)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    return b200::launch_ew(ctx, 1, cols, [=] __device__(int64_t, int64_t j) { stop_status[j] = 0; });
}

#define B200_DEF_STEPS(V, VT)                                                                  \
    b200_status b200_cg_initialize_##V(b200_ctx* ctx, int64_t rows, int64_t cols, const VT* b, \
                                       int64_t bs, VT* r, int64_t rs, VT* z, int64_t zs,       \
                                       VT* p, int64_t ps, VT* q, int64_t qs, VT* prev_rho,     \
                                       VT* rho, uint8_t* stop)                                 \
    {                                                                                          \
        return b200::steps::cg_initialize<VT>(ctx, rows, cols, b, bs, r, rs, z, zs, p, ps, q, qs,     \
                                       prev_rho, rho, stop);                                   \
    }                                                                                          \
    b200_status b200_cg_step_1_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* p,           \
                                   int64_t ps, const VT* z, int64_t zs, const VT* rho,         \
                                   const VT* prev_rho, const uint8_t* stop)                    \
    {                                                                                          \
        return b200::steps::cg_step_1<VT>(ctx, rows, cols, p, ps, z, zs, rho, prev_rho, stop);        \
    }                                                                                          \
    b200_status b200_cg_step_2_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* x,           \
                                   int64_t xs, VT* r, int64_t rs, const VT* p, int64_t ps,     \
                                   const VT* q, int64_t qs, const VT* beta, const VT* rho,     \
                                   const uint8_t* stop)                                        \
    {                                                                                          \
        return b200::steps::cg_step_2<VT>(ctx, rows, cols, x, xs, r, rs, p, ps, q, qs, beta, rho,     \
                                   stop);                                                      \
    }                                                                                          \
    b200_status b200_fcg_initialize_##V(                                                       \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* b, int64_t bs, VT* r, int64_t rs, \
        VT* z, int64_t zs, VT* p, int64_t ps, VT* q, int64_t qs, VT* t, int64_t ts,            \
        VT* prev_rho, VT* rho, VT* rho_t, uint8_t* stop)                                       \
    {                                                                                          \
        return b200::steps::fcg_initialize<VT>(ctx, rows, cols, b, bs, r, rs, z, zs, p, ps, q, \
                                               qs, t, ts, prev_rho, rho, rho_t, stop);         \
    }                                                                                          \
    b200_status b200_fcg_step_1_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* p,          \
                                    int64_t ps, const VT* z, int64_t zs, const VT* rho_t,      \
                                    const VT* prev_rho, const uint8_t* stop)                   \
    {                                                                                          \
        return b200::steps::fcg_step_1<VT>(ctx, rows, cols, p, ps, z, zs, rho_t, prev_rho,     \
                                           stop);                                              \
    }                                                                                          \
    b200_status b200_fcg_step_2_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* x,          \
                                    int64_t xs, VT* r, int64_t rs, VT* t, int64_t ts,          \
                                    const VT* p, int64_t ps, const VT* q, int64_t qs,          \
                                    const VT* beta, const VT* rho, const uint8_t* stop)        \
    {                                                                                          \
        return b200::steps::fcg_step_2<VT>(ctx, rows, cols, x, xs, r, rs, t, ts, p, ps, q, qs, \
                                           beta, rho, stop);                                   \
    }                                                                                          \
    b200_status b200_cgs_initialize_##V(                                                       \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* b, int64_t bs, VT* r, int64_t rs, \
        VT* r_tld, int64_t rts, VT* p, int64_t ps, VT* q, int64_t qs, VT* u, int64_t us,       \
        VT* u_hat, int64_t uhs, VT* v_hat, int64_t vhs, VT* t, int64_t ts, VT* alpha,          \
        VT* beta, VT* gamma, VT* prev_rho, VT* rho, uint8_t* stop)                             \
    {                                                                                          \
        return b200::steps::cgs_initialize<VT>(ctx, rows, cols, b, bs, r, rs, r_tld, rts, p,   \
                                               ps, q, qs, u, us, u_hat, uhs, v_hat, vhs, t,    \
                                               ts, alpha, beta, gamma, prev_rho, rho, stop);   \
    }                                                                                          \
    b200_status b200_cgs_step_1_##V(b200_ctx* ctx, int64_t rows, int64_t cols, const VT* r,    \
                                    int64_t rs, VT* u, int64_t us, VT* p, int64_t ps,          \
                                    const VT* q, int64_t qs, VT* beta, const VT* rho,          \
                                    const VT* prev_rho, const uint8_t* stop)                   \
    {                                                                                          \
        return b200::steps::cgs_step_1<VT>(ctx, rows, cols, r, rs, u, us, p, ps, q, qs, beta,  \
                                           rho, prev_rho, stop);                               \
    }                                                                                          \
    b200_status b200_cgs_step_2_##V(b200_ctx* ctx, int64_t rows, int64_t cols, const VT* u,    \
                                    int64_t us, const VT* v_hat, int64_t vhs, VT* q,           \
                                    int64_t qs, VT* t, int64_t ts, VT* alpha, const VT* rho,   \
                                    const VT* gamma, const uint8_t* stop)                      \
    {                                                                                          \
        return b200::steps::cgs_step_2<VT>(ctx, rows, cols, u, us, v_hat, vhs, q, qs, t, ts,   \
                                           alpha, rho, gamma, stop);                           \
    }                                                                                          \
    b200_status b200_cgs_step_3_##V(b200_ctx* ctx, int64_t rows, int64_t cols, const VT* t,    \
                                    int64_t ts, const VT* u_hat, int64_t uhs, VT* r,           \
                                    int64_t rs, VT* x, int64_t xs, const VT* alpha,            \
                                    const uint8_t* stop)                                       \
    {                                                                                          \
        return b200::steps::cgs_step_3<VT>(ctx, rows, cols, t, ts, u_hat, uhs, r, rs, x, xs,   \
                                           alpha, stop);                                       \
    }                                                                                          \
    b200_status b200_chebyshev_init_update_##V(b200_ctx* ctx, int64_t rows, int64_t cols,      \
                                               double alpha, const VT* inner_sol, int64_t is,  \
                                               VT* update_sol, int64_t us, VT* output,         \
                                               int64_t os)                                     \
    {                                                                                          \
        return b200::steps::chebyshev_init_update<VT>(ctx, rows, cols, alpha, inner_sol, is,   \
                                                      update_sol, us, output, os);             \
    }                                                                                          \
    b200_status b200_chebyshev_update_##V(b200_ctx* ctx, int64_t rows, int64_t cols,           \
                                          double alpha, double beta, VT* inner_sol,            \
                                          int64_t is, VT* update_sol, int64_t us, VT* output,  \
                                          int64_t os)                                          \
    {                                                                                          \
        return b200::steps::chebyshev_update<VT>(ctx, rows, cols, alpha, beta, inner_sol, is,  \
                                                 update_sol, us, output, os);                  \
    }                                                                                          \
    b200_status b200_pipe_cg_initialize_1_##V(b200_ctx* ctx, int64_t rows, int64_t cols,       \
                                              const VT* b, int64_t bs, VT* r, int64_t rs,      \
                                              VT* prev_rho, uint8_t* stop)                     \
    {                                                                                          \
        return b200::steps::pipe_cg_initialize_1<VT>(ctx, rows, cols, b, bs, r, rs, prev_rho,  \
                                                     stop);                                    \
    }                                                                                          \
    b200_status b200_pipe_cg_initialize_2_##V(                                                 \
        b200_ctx* ctx, int64_t rows, int64_t cols, VT* p, int64_t ps, VT* q, int64_t qs,       \
        VT* f, int64_t fs, VT* g, int64_t gs, VT* beta, const VT* z, int64_t zs, const VT* w,  \
        int64_t ws, const VT* m, int64_t ms, const VT* n, int64_t ns, const VT* delta)         \
    {                                                                                          \
        return b200::steps::pipe_cg_initialize_2<VT>(ctx, rows, cols, p, ps, q, qs, f, fs, g,  \
                                                     gs, beta, z, zs, w, ws, m, ms, n, ns,     \
                                                     delta);                                   \
    }                                                                                          \
    b200_status b200_pipe_cg_step_1_##V(                                                       \
        b200_ctx* ctx, int64_t rows, int64_t cols, VT* x, int64_t xs, VT* r, int64_t rs,       \
        VT* z1, int64_t z1s, VT* z2, int64_t z2s, VT* w, int64_t ws, const VT* p, int64_t ps,  \
        const VT* q, int64_t qs, const VT* f, int64_t fs, const VT* g, int64_t gs,             \
        const VT* rho, const VT* beta, const uint8_t* stop)                                    \
    {                                                                                          \
        return b200::steps::pipe_cg_step_1<VT>(ctx, rows, cols, x, xs, r, rs, z1, z1s, z2,     \
                                               z2s, w, ws, p, ps, q, qs, f, fs, g, gs, rho,    \
                                               beta, stop);                                    \
    }                                                                                          \
    b200_status b200_pipe_cg_step_2_##V(                                                       \
        b200_ctx* ctx, int64_t rows, int64_t cols, VT* beta, VT* p, int64_t ps, VT* q,         \
        int64_t qs, VT* f, int64_t fs, VT* g, int64_t gs, const VT* z, int64_t zs,             \
        const VT* w, int64_t ws, const VT* m, int64_t ms, const VT* n, int64_t ns,             \
        const VT* prev_rho, const VT* rho, const VT* delta, const uint8_t* stop)               \
    {                                                                                          \
        return b200::steps::pipe_cg_step_2<VT>(ctx, rows, cols, beta, p, ps, q, qs, f, fs, g,  \
                                               gs, z, zs, w, ws, m, ms, n, ns, prev_rho, rho,  \
                                               delta, stop);                                   \
    }                                                                                          \
    b200_status b200_gcr_initialize_##V(b200_ctx* ctx, int64_t rows, int64_t cols,             \
                                        const VT* b, int64_t bs, VT* residual, int64_t rs,     \
                                        uint8_t* stop)                                         \
    {                                                                                          \
        return b200::steps::gcr_initialize<VT>(ctx, rows, cols, b, bs, residual, rs, stop);    \
    }                                                                                          \
    b200_status b200_gcr_restart_##V(b200_ctx* ctx, int64_t rows, int64_t cols,                \
                                     const VT* residual, int64_t rs, const VT* a_residual,     \
                                     int64_t ars, VT* p_bases, int64_t ps, VT* ap_bases,       \
                                     int64_t aps, uint64_t* final_iter_nums)                   \
    {                                                                                          \
        return b200::steps::gcr_restart<VT>(ctx, rows, cols, residual, rs, a_residual, ars,    \
                                            p_bases, ps, ap_bases, aps, final_iter_nums);      \
    }                                                                                          \
    b200_status b200_gcr_step_1_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* x,          \
                                    int64_t xs, VT* residual, int64_t rs, const VT* p,         \
                                    int64_t ps, const VT* ap, int64_t aps, const VT* ap_norm,  \
                                    const VT* rap, const uint8_t* stop)                        \
    {                                                                                          \
        return b200::steps::gcr_step_1<VT>(ctx, rows, cols, x, xs, residual, rs, p, ps, ap,    \
                                           aps, ap_norm, rap, stop);                           \
    }                                                                                          \
    b200_status b200_minres_initialize_##V(                                                    \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* r, int64_t rs, VT* z, int64_t zs, \
        VT* p, int64_t ps, VT* p_prev, int64_t pps, VT* q, int64_t qs, VT* q_prev,             \
        int64_t qps, VT* q_tilde, int64_t qts, VT* beta, VT* gamma, VT* delta, VT* cos_prev,   \
        VT* cos_, VT* sin_prev, VT* sin_, VT* eta_next, VT* eta, uint8_t* stop)                \
    {                                                                                          \
        return b200::steps::minres_initialize<VT>(ctx, rows, cols, r, rs, z, zs, p, ps,        \
                                                  p_prev, pps, q, qs, q_prev, qps, q_tilde,    \
                                                  qts, beta, gamma, delta, cos_prev, cos_,     \
                                                  sin_prev, sin_, eta_next, eta, stop);        \
    }                                                                                          \
    b200_status b200_minres_step_1_##V(b200_ctx* ctx, int64_t cols, VT* alpha, VT* beta,       \
                                       VT* gamma, VT* delta, VT* cos_prev, VT* cos_,           \
                                       VT* sin_prev, VT* sin_, VT* eta, VT* eta_next,          \
                                       VT* tau, const uint8_t* stop)                           \
    {                                                                                          \
        return b200::steps::minres_step_1<VT>(ctx, cols, alpha, beta, gamma, delta, cos_prev,  \
                                              cos_, sin_prev, sin_, eta, eta_next, tau, stop); \
    }                                                                                          \
    b200_status b200_minres_step_2_##V(                                                        \
        b200_ctx* ctx, int64_t rows, int64_t cols, VT* x, int64_t xs, VT* p, int64_t ps,       \
        const VT* p_prev, int64_t pps, VT* z, int64_t zs, const VT* z_tilde, int64_t zts,      \
        VT* q, int64_t qs, VT* q_prev, int64_t qps, VT* v, int64_t vs, const VT* alpha,        \
        const VT* beta, const VT* gamma, const VT* delta, const VT* cos_, const VT* eta,       \
        const uint8_t* stop)                                                                   \
    {                                                                                          \
        return b200::steps::minres_step_2<VT>(ctx, rows, cols, x, xs, p, ps, p_prev, pps, z,   \
                                              zs, z_tilde, zts, q, qs, q_prev, qps, v, vs,     \
                                              alpha, beta, gamma, delta, cos_, eta, stop);     \
    }                                                                                          \
    b200_status b200_bicgstab_initialize_##V(                                                  \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* b, int64_t bs, VT* r, int64_t rs, \
        VT* rr, int64_t rrs, VT* y, int64_t ys, VT* s, int64_t ss, VT* t, int64_t ts, VT* z,   \
        int64_t zs, VT* v, int64_t vs, VT* p, int64_t ps, VT* prev_rho, VT* rho, VT* alpha,    \
        VT* beta, VT* gamma, VT* omega, uint8_t* stop)                                         \
    {                                                                                          \
        return b200::steps::bicgstab_initialize<VT>(ctx, rows, cols, b, bs, r, rs, rr, rrs, y, ys, s, \
                                             ss, t, ts, z, zs, v, vs, p, ps, prev_rho, rho,    \
                                             alpha, beta, gamma, omega, stop);                 \
    }                                                                                          \
    b200_status b200_bicgstab_step_1_##V(                                                      \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* r, int64_t rs, VT* p, int64_t ps, \
        const VT* v, int64_t vs, const VT* rho, const VT* prev_rho, const VT* alpha,           \
        const VT* omega, const uint8_t* stop)                                                  \
    {                                                                                          \
        return b200::steps::bicgstab_step_1<VT>(ctx, rows, cols, r, rs, p, ps, v, vs, rho, prev_rho,  \
                                         alpha, omega, stop);                                  \
    }                                                                                          \
    b200_status b200_bicgstab_step_2_##V(                                                      \
        b200_ctx* ctx, int64_t rows, int64_t cols, const VT* r, int64_t rs, VT* s, int64_t ss, \
        const VT* v, int64_t vs, const VT* rho, VT* alpha, const VT* beta,                     \
        const uint8_t* stop)                                                                   \
    {                                                                                          \
        return b200::steps::bicgstab_step_2<VT>(ctx, rows, cols, r, rs, s, ss, v, vs, rho, alpha,     \
                                         beta, stop);                                          \
    }                                                                                          \
    b200_status b200_bicgstab_step_3_##V(                                                      \
        b200_ctx* ctx, int64_t rows, int64_t cols, VT* x, int64_t xs, VT* r, int64_t rs,       \
        const VT* s, int64_t ss, const VT* t, int64_t ts, const VT* y, int64_t ys,             \
        const VT* z, int64_t zs, const VT* alpha, const VT* beta, const VT* gamma, VT* omega,  \
        const uint8_t* stop)                                                                   \
    {                                                                                          \
        return b200::steps::bicgstab_step_3<VT>(ctx, rows, cols, x, xs, r, rs, s, ss, t, ts, y, ys,   \
                                         z, zs, alpha, beta, gamma, omega, stop);              \
    }                                                                                          \
    b200_status b200_bicgstab_finalize_##V(b200_ctx* ctx, int64_t rows, int64_t cols, VT* x,   \
                                           int64_t xs, const VT* y, int64_t ys,                \
                                           const VT* alpha, uint8_t* stop)                     \
    {                                                                                          \
        return b200::steps::bicgstab_finalize<VT>(ctx, rows, cols, x, xs, y, ys, alpha, stop);        \
    }

B200_DEF_STEPS(f64, double)
B200_DEF_STEPS(f32, float)

}  // extern "C"
