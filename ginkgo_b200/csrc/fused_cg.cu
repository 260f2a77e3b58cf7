// Fused CG iteration for sm_100a: alpha / beta / rho are produced AND consumed on
// the device, three launches per iteration, no host round trip.
//
// The reference drives CG from the host with 8 launches and a blocking
// device->host copy per iteration (core/solver/cg.cpp:142-180: preconditioner
// apply, dot, norm2 + stopping check, step_1, SpMV, dot, step_2).  The same
// recurrence, regrouped around its two unavoidable global reductions:
//
//   step_p   p = z + (rho/prev_rho) p                                  reads 2n, writes n
//   spmv_dot q = A p  and  pq = p.q   (fused in the SpMV epilogue)     matrix + 2n
//   step_xr  alpha = rho/pq; x += alpha p; r -= alpha q; z = M^-1 r;
//            rho' = r.z; rr = r.r; stopping check (in the last CTA)    reads 5n, writes 3n
//
// = matrix + 13 n values per iteration, against 19 n for
// the reference's kernel sequence.  The stopping check is the reference's
// (core/stop/{combined,iteration,residual_norm}.cpp): Iteration and
// (Implicit)ResidualNorm in the user's order, ids 1 and 2, written into a
// device control block; once ctl[0] != 0 every kernel of later iterations is a
// no-op, so a CUDA graph of k iterations can be replayed blindly and the host
// only reads the control block every k iterations.
//
// Scalars  sc[8]: 0 rho, 1 prev_rho, 2 pq, 3 rr, 4 tau0 (baseline norm),
//                 5 threshold = reduction_factor * tau0, 6/7 local partial sums
//                 (multi-GPU: all-reduced between step_xr and finish)
// Control  ctl[8] (int32): 0 stopping_status byte (0 = running), 1 iterations done,
//                 2 max_iters (-1 none), 3 res_kind (0 none, 1 ResidualNorm,
//                 2 ImplicitResidualNorm), 4 iter_first
#include "csr_launch.cuh"

namespace b200 {
namespace fcg {

constexpr int kThreads = 512;

template <typename V>
__global__ void __launch_bounds__(kThreads)
    step_p_kernel(int64_t n, V* __restrict__ p, const V* __restrict__ z, const V* __restrict__ sc,
                  const int32_t* __restrict__ ctl)
{
    if (ctl[0] != 0) return;
    const V rho = sc[0], prev = sc[1];
    const bool plain = (prev == V(0));
    const V beta = plain ? V(0) : rho / prev;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const V zi = z[i];
        p[i] = plain ? zi : zi + beta * p[i];
    }
}

// the reference's criteria, evaluated by ONE thread (core/stop/combined.cpp:33-52)
template <typename V>
__device__ void finish_scalars(V rho_new, V rr, V* sc, int32_t* ctl)
{
    sc[1] = sc[0];
    sc[0] = rho_new;
    sc[3] = rr;
    const int32_t iter = ctl[1] + 1;
    ctl[1] = iter;
    const int32_t max_iters = ctl[2], res_kind = ctl[3], iter_first = ctl[4];
    const bool has_it = max_iters >= 0, has_res = res_kind != 0;
    uint8_t status = 0;
    uint8_t id = 1;
    for (int k = 0; k < 2 && status == 0; ++k) {
        const bool is_it = (k == 0) == (iter_first != 0);
        if (is_it) {
            if (!has_it) continue;
            if (iter >= max_iters) status = (id & kIdMask) | kFinalizedMask;
            ++id;
        } else {
            if (!has_res) continue;
            const V tau = res_kind == 1 ? sqrt(rr) : sqrt(fabs(rho_new));
            if (tau <= sc[5]) status = kConvergedMask | (id & kIdMask) | kFinalizedMask;
            ++id;
        }
    }
    ctl[0] = status;
}

// INIT: r already holds b - A x; computes z, rho, rr (no x/r update) and the baseline.
template <typename V, bool INIT>
__global__ void __launch_bounds__(kThreads)
    step_xr_kernel(int64_t n, V* __restrict__ x, V* __restrict__ r, const V* __restrict__ p,
                   const V* __restrict__ q, V* __restrict__ z, const V* __restrict__ inv_diag,
                   V* __restrict__ sc, int32_t* __restrict__ ctl, V* __restrict__ partials,
                   unsigned int* __restrict__ counter, int finalize, int baseline, V factor)
{
    __shared__ V red[32];
    __shared__ bool is_last;
    if (!INIT && ctl[0] != 0) return;
    const int tid = threadIdx.x;
    const V pq = sc[2];
    const bool update = !INIT && pq != V(0);
    const V alpha = update ? sc[0] / pq : V(0);
    V a_rho = V(0), a_rr = V(0);
    for (int64_t i = blockIdx.x * (int64_t)kThreads + tid; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        V ri = r[i];
        if (update) {
            x[i] += alpha * p[i];
            ri -= alpha * q[i];
            r[i] = ri;
        }
        const V zi = inv_diag ? ri * inv_diag[i] : ri;
        z[i] = zi;
        a_rho += ri * zi;
        a_rr += ri * ri;
    }
    a_rho = block_sum(a_rho, red);
    a_rr = block_sum(a_rr, red);
    if (tid == 0) {
        partials[2 * blockIdx.x] = a_rho;
        partials[2 * blockIdx.x + 1] = a_rr;
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    V s_rho = V(0), s_rr = V(0);
    for (int k = tid; k < (int)gridDim.x; k += kThreads) {
        s_rho += __ldcg(partials + 2 * k);
        s_rr += __ldcg(partials + 2 * k + 1);
    }
    s_rho = block_sum(s_rho, red);
    s_rr = block_sum(s_rr, red);
    if (tid == 0) {
        *counter = 0u;
        if (!finalize) {
            sc[6] = s_rho;
            sc[7] = s_rr;
        } else {
            if (INIT) {
                // baseline: 0 rhs_norm (sc[4] preset to ||b||), 1 initial_resnorm, 2 absolute
                if (baseline == 1) sc[4] = sqrt(s_rr);
                if (baseline == 2) sc[4] = V(1);
                sc[5] = factor * sc[4];
            }
            finish_scalars(s_rho, s_rr, sc, ctl);
        }
    }
}

// multi-GPU: after the all-reduce of sc[6..7] (and of sc[2] for pq)
template <typename V, bool INIT>
__global__ void finish_kernel(V* sc, int32_t* ctl, int baseline, V factor)
{
    if (!INIT && ctl[0] != 0) return;
    if (INIT) {
        // baseline: 1 initial_resnorm, 2 absolute, 3 rhs_norm with sc[4] = global ||b||^2,
        // 0 rhs_norm with sc[4] = ||b|| already
        if (baseline == 1) sc[4] = sqrt(sc[7]);
        if (baseline == 2) sc[4] = V(1);
        if (baseline == 3) sc[4] = sqrt(sc[4]);
        sc[5] = factor * sc[4];
    }
    finish_scalars(sc[6], sc[7], sc, ctl);
}

template <typename V>
__global__ void init_scalars_kernel(V* sc, int32_t* ctl, int32_t max_iters, int32_t res_kind,
                                    int32_t iter_first)
{
    sc[0] = V(1);  // becomes prev_rho = 1 after the INIT finish (reference cg::initialize)
    sc[1] = V(1);
    sc[2] = V(0);
    sc[3] = V(0);
    sc[6] = V(0);
    sc[7] = V(0);
    ctl[0] = 0;
    ctl[1] = -1;  // the INIT finish increments to 0
    ctl[2] = max_iters;
    ctl[3] = res_kind;
    ctl[4] = iter_first;
}

inline int ew_grid(const b200_ctx* ctx, int64_t n)
{
    int64_t g = ceildiv(n, kThreads * 4);
    const int64_t cap = (int64_t)ctx->num_sms * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace fcg
}  // namespace b200

struct b200_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

extern "C" {

// ---- CUDA-graph helpers: capture a sequence of C-ABI calls on the context's stream -------
b200_status b200_graph_begin_capture(b200_ctx* ctx)
{
    B200_REQUIRE(ctx != nullptr, "ctx is null");
    B200_CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    return B200_OK;
}

b200_status b200_graph_end_capture(b200_ctx* ctx, b200_graph** out)
{
    B200_REQUIRE(ctx && out, "null argument");
    b200_graph* g = new b200_graph();
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
    if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        b200::set_error("graph capture failed: %s", cudaGetErrorString(e));
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
        cudaGetLastError();
        return B200_ERR_CUDA;
    }
    *out = g;
    return B200_OK;
}

b200_status b200_graph_launch(b200_ctx* ctx, b200_graph* graph)
{
    B200_REQUIRE(ctx && graph, "null argument");
    B200_CUDA_CHECK(cudaGraphLaunch(graph->exec, ctx->stream));
    return B200_OK;
}

void b200_graph_destroy(b200_graph* graph)
{
    if (!graph) return;
    if (graph->exec) cudaGraphExecDestroy(graph->exec);
    if (graph->graph) cudaGraphDestroy(graph->graph);
    delete graph;
}

#define B200_DEF_FCG(V, VT)                                                                      \
    /* elements of `work` (value type) the fused kernels need for a matrix of this size */       \
    int64_t b200_cg_fused_work_size_##V(const b200_ctx* ctx)                                     \
    {                                                                                            \
        return (int64_t)ctx->num_sms * 4 + 2 * (int64_t)ctx->num_sms * 4 + 64;                   \
    }                                                                                            \
    b200_status b200_cg_fused_init_##V(                                                          \
        b200_ctx* ctx, int64_t n, const VT* r, VT* z, VT* p, VT* q, const VT* inv_diag, VT* sc,  \
        int32_t* ctl, VT* work, int64_t max_iters, int32_t res_kind, int32_t iter_first,         \
        int32_t baseline, VT reduction_factor, int32_t finalize)                                 \
    {                                                                                            \
        B200_REQUIRE(ctx && sc && ctl && work, "null argument");                                 \
        b200::fcg::init_scalars_kernel<VT><<<1, 1, 0, ctx->stream>>>(                            \
            sc, ctl, (int32_t)(max_iters > 0x7fffffff ? 0x7fffffff : max_iters), res_kind,       \
            iter_first);                                                                         \
        B200_LAUNCH_CHECK(ctx);                                                                  \
        B200_CUDA_CHECK(cudaMemsetAsync(p, 0, sizeof(VT) * n, ctx->stream));                     \
        B200_CUDA_CHECK(cudaMemsetAsync(q, 0, sizeof(VT) * n, ctx->stream));                     \
        const int grid = b200::fcg::ew_grid(ctx, n);                                             \
        b200::fcg::step_xr_kernel<VT, true><<<grid, b200::fcg::kThreads, 0, ctx->stream>>>(      \
            n, nullptr, const_cast<VT*>(r), p, q, z, inv_diag, sc, ctl,                          \
            work + (int64_t)ctx->num_sms * 4, ctx->counters + 2, finalize, baseline,             \
            reduction_factor);                                                                   \
        B200_LAUNCH_CHECK(ctx);                                                                  \
        return B200_OK;                                                                          \
    }                                                                                            \
    b200_status b200_cg_fused_step_p_##V(b200_ctx* ctx, int64_t n, VT* p, const VT* z,           \
                                         const VT* sc, const int32_t* ctl)                       \
    {                                                                                            \
        const int grid = b200::fcg::ew_grid(ctx, n);                                             \
        b200::fcg::step_p_kernel<VT><<<grid, b200::fcg::kThreads, 0, ctx->stream>>>(n, p, z, sc, \
                                                                                    ctl);        \
        B200_LAUNCH_CHECK(ctx);                                                                  \
        return B200_OK;                                                                          \
    }                                                                                            \
    b200_status b200_cg_fused_step_xr_##V(b200_ctx* ctx, int64_t n, VT* x, VT* r, const VT* p,   \
                                          const VT* q, VT* z, const VT* inv_diag, VT* sc,        \
                                          int32_t* ctl, VT* work, int32_t finalize)              \
    {                                                                                            \
        const int grid = b200::fcg::ew_grid(ctx, n);                                             \
        b200::fcg::step_xr_kernel<VT, false><<<grid, b200::fcg::kThreads, 0, ctx->stream>>>(     \
            n, x, r, p, q, z, inv_diag, sc, ctl, work + (int64_t)ctx->num_sms * 4,               \
            ctx->counters + 2, finalize, 0, VT(0));                                              \
        B200_LAUNCH_CHECK(ctx);                                                                  \
        return B200_OK;                                                                          \
    }                                                                                            \
    b200_status b200_cg_fused_finish_##V(b200_ctx* ctx, VT* sc, int32_t* ctl, int32_t init,      \
                                         int32_t baseline, VT reduction_factor)                  \
    {                                                                                            \
        if (init)                                                                                \
            b200::fcg::finish_kernel<VT, true><<<1, 1, 0, ctx->stream>>>(sc, ctl, baseline,      \
                                                                         reduction_factor);      \
        else                                                                                     \
            b200::fcg::finish_kernel<VT, false><<<1, 1, 0, ctx->stream>>>(sc, ctl, baseline,     \
                                                                          reduction_factor);     \
        B200_LAUNCH_CHECK(ctx);                                                                  \
        return B200_OK;                                                                          \
    }

B200_DEF_FCG(f64, double)
B200_DEF_FCG(f32, float)

/* c = A b and *dot_out = b . c in one launch (square A; b is also the dot operand).  `work`  */
/* provides the per-CTA partials (b200_cg_fused_work_size elements); ctl may be NULL.         */
#define B200_DEF_SPMV_DOT(V, VT, I, IT)                                                          \
    b200_status b200_csr_spmv_dot_##V##_##I(                                                     \
        b200_ctx* ctx, const b200_csr_plan* plan, int64_t num_rows, int64_t num_cols,            \
        int64_t nnz, const IT* row_ptrs, const IT* col_idxs, const VT* values, const VT* b,      \
        VT* c, VT* dot_out, VT* work, const int32_t* ctl)                                        \
    {                                                                                            \
        B200_REQUIRE(ctx && plan && dot_out && work, "null argument (a plan is required)");      \
        B200_REQUIRE(plan->num_rows == num_rows && plan->nnz == nnz, "plan does not match");     \
        const b200::csr::Variant v = b200::csr::pick_variant(col_idxs, values, plan);            \
        B200_REQUIRE(v != b200::csr::kSlab, "col_idxs/values must be 32-byte aligned");          \
        b200::csr::DotArgs<VT> dot{work, ctx->counters + 1, dot_out, ctl};                       \
        return b200::csr::launch_planned<VT, IT, false, true>(                                   \
            ctx, plan, v, nnz, row_ptrs, col_idxs, values, nullptr, b, 1, nullptr, c, 1, dot);   \
    }

B200_DEF_SPMV_DOT(f64, double, i32, int32_t)
B200_DEF_SPMV_DOT(f64, double, i64, int64_t)
B200_DEF_SPMV_DOT(f32, float, i32, int32_t)
B200_DEF_SPMV_DOT(f32, float, i64, int64_t)

}  // extern "C"
