// gko_b200_dist.hpp -- multi-GPU host layer: the NCCL counterpart of the reference's
// experimental::distributed::{Matrix, Vector} + distributed CG on this path
// (core/distributed/matrix.cpp:450-509, core/distributed/vector.cpp:510-534,
// core/solver/cg.cpp driven through precision_dispatch_real_complex_distributed).
// One process per GPU; rank p owns rows [offsets[p], offsets[p+1]).
#pragma once

#include <cstdio>
#include <cstdlib>

#include "gko_b200.hpp"

namespace gko_b200 {
namespace distributed {

// experimental::mpi::communicator analogue on NCCL
class communicator {
public:
    static void get_unique_id(uint8 (&id)[128]) { GKOB_CALL(b200_comm_get_unique_id(id)); }
    static std::shared_ptr<communicator> create(std::shared_ptr<const Executor> exec,
                                                const uint8* id128, int rank, int size)
    {
        auto c = std::shared_ptr<communicator>(new communicator());
        c->exec_ = exec;
        GKOB_CALL(b200_comm_create(exec->ctx(), id128, rank, size, &c->comm_));
        // peer-memory collectives over NVLink/NVSwitch unless B200_P2P=0; when CUDA IPC is not
        // available between the ranks every rank gets the same error and NCCL stays in use
        const char* env = std::getenv("B200_P2P");
        c->want_p2p_ = !(env && env[0] == '0');
        if (c->want_p2p_ && size > 1 &&
            b200_comm_enable_p2p(exec->ctx(), c->comm_) != B200_OK) {
            std::fprintf(stderr, "[gko_b200] rank %d: peer-memory collectives unavailable (%s), using NCCL\n",
                         rank, b200_last_error());
            c->want_p2p_ = false;
        }
        return c;
    }
    bool use_p2p() const { return want_p2p_ && b200_comm_p2p_enabled(comm_); }
    // throws if a peer-memory wait timed out (synchronises the stream)
    void check() const
    {
        if (b200_comm_p2p_error(exec_->ctx(), comm_))
            throw Error("peer-memory collective timed out: a rank stopped or the call sequences diverged");
    }
    ~communicator() { b200_comm_destroy(comm_); }
    int rank() const { return b200_comm_rank(comm_); }
    int size() const { return b200_comm_size(comm_); }
    b200_comm* get() const { return comm_; }

private:
    communicator() = default;
    std::shared_ptr<const Executor> exec_;
    b200_comm* comm_ = nullptr;
    bool want_p2p_ = true;
};

template <typename V>
struct cabi;
template <>
struct cabi<double> {
    static constexpr auto allreduce = b200_comm_allreduce_sum_f64;
    static constexpr auto halo_exchange = b200_halo_exchange_f64;
};
template <>
struct cabi<float> {
    static constexpr auto allreduce = b200_comm_allreduce_sum_f32;
    static constexpr auto halo_exchange = b200_halo_exchange_f32;
};

// distributed::Matrix: local rows, columns numbered into the extended vector
// [n_local owned | n_ghost received]; apply = halo exchange + one local SpMV.
template <typename V, typename I>
class Matrix {
public:
    Matrix(std::shared_ptr<const Executor> exec, std::shared_ptr<communicator> comm,
           std::unique_ptr<matrix::Csr<V, I>> local, size_type n_ghost,
           const std::vector<int64>& send_counts, const std::vector<int64>& recv_counts,
           const int32* send_idx_dev)
        : exec_(exec), comm_(comm), local_(std::move(local)), n_ghost_(n_ghost)
    {
        if (local_->get_size().cols != local_->get_size().rows + n_ghost)
            throw BadDimension("distributed::Matrix: local block must be n_local x (n_local+n_ghost)");
        GKOB_CALL(b200_halo_create(exec->ctx(), comm->size(), local_->get_size().rows, n_ghost,
                                   send_counts.data(), recv_counts.data(), send_idx_dev,
                                   (int32)sizeof(V), &halo_));
        if (comm->use_p2p() &&
            b200_halo_enable_p2p(exec->ctx(), comm->get(), halo_) != B200_OK)
            std::fprintf(stderr, "[gko_b200] rank %d: peer-memory halo unavailable (%s), using NCCL\n",
                         comm->rank(), b200_last_error());
    }
    ~Matrix() { b200_halo_destroy(halo_); }
    size_type n_local() const { return local_->get_size().rows; }
    size_type n_ghost() const { return n_ghost_; }
    const matrix::Csr<V, I>* get_local_matrix() const { return local_.get(); }
    std::shared_ptr<communicator> get_communicator() const { return comm_; }
    b200_halo* get_halo() const { return halo_; }
    // y_local = A x   (x_ext: owned part filled by the caller, ghosts by the exchange)
    void apply(matrix::Dense<V>* x_ext, matrix::Dense<V>* y_local) const
    {
        GKOB_CALL(cabi<V>::halo_exchange(exec_->ctx(), comm_->get(), halo_, x_ext->get_values(),
                                         nullptr));
        local_->apply(x_ext, y_local);
    }

private:
    std::shared_ptr<const Executor> exec_;
    std::shared_ptr<communicator> comm_;
    std::unique_ptr<matrix::Csr<V, I>> local_;
    size_type n_ghost_;
    b200_halo* halo_ = nullptr;
};

// Distributed CG with the fused device-resident iteration: per iteration
//   step_p | halo exchange of p | spmv_dot | all-reduce(pq) | step_xr | all-reduce(rho, rr) | finish
// all enqueued on one stream and captured in a CUDA graph of `check_every` iterations.
template <typename V, typename I>
class Cg {
public:
    using Dense = matrix::Dense<V>;
    Cg(std::shared_ptr<const Executor> exec, std::shared_ptr<const Matrix<V, I>> A,
       bool scalar_jacobi, int64 max_iters, int res_kind, int baseline, double reduction,
       bool iter_first, int check_every)
        : exec_(exec), A_(A), max_iters_(max_iters), res_kind_(res_kind), baseline_(baseline),
          reduction_(reduction), iter_first_(iter_first), check_every_(std::max(1, check_every))
    {
        const size_type n = A->n_local();
        if (scalar_jacobi) {
            // the diagonal is local: column index == row index in the extended numbering
            auto d = A->get_local_matrix()->extract_diagonal();
            inv_diag_ = array<V>(exec, n);
            GKOB_CALL(vabi<V>::invert_diagonal(exec->ctx(), n, d->get_const_values(),
                                               inv_diag_.get_data()));
        }
        ws_ = Dense::create(exec, dim2{3 * n + 2 * (n + A->n_ghost()), 1});
        sc_ = array<V>(exec, 8);
        ctl_ = array<int32>(exec, 8);
        work_ = array<V>(exec, (size_type)vabi<V>::fused_work_size(exec->ctx()));
        one_ = matrix::scalar<V>(V(1), exec);
        neg_one_ = matrix::scalar<V>(V(-1), exec);
    }
    ~Cg() { b200_graph_destroy(graph_); }

    // b_local, x_local: this rank's n_local entries
    void apply(const Dense* b, Dense* x)
    {
        auto ctx = exec_->ctx();
        auto comm = A_->get_communicator()->get();
        const int64 n = A_->n_local(), ng = A_->n_ghost();
        auto Al = A_->get_local_matrix();
        const int64 nnz = Al->get_num_stored_elements();
        V* r = ws_->get_values();
        V* z = r + n;
        V* q = z + n;
        V* p_ext = q + n;
        V* x_ext = p_ext + (n + ng);
        const V* dinv = inv_diag_.get_size() ? inv_diag_.get_const_data() : nullptr;
        V* xv = x->get_values();
        V* sc = sc_.get_data();
        int32* ctl = ctl_.get_data();
        // r = b - A x
        exec_->copy(x_ext, x->get_const_values(), n);
        exec_->copy(r, b->get_const_values(), n);
        auto xe = Dense::create_view(exec_, dim2{(size_type)(n + ng), 1}, x_ext, 1);
        auto rv = Dense::create_view(exec_, dim2{(size_type)n, 1}, r, 1);
        GKOB_CALL(cabi<V>::halo_exchange(ctx, comm, A_->get_halo(), x_ext, nullptr));
        Al->apply(neg_one_.get(), xe.get(), one_.get(), rv.get());
        // global ||b||^2 -> sc[4] (the INIT finish takes the square root for rhs_norm)
        GKOB_CALL(vabi<V>::sqnorm2(ctx, n, 1, b->get_const_values(), 1, sc + 4));
        GKOB_CALL(cabi<V>::allreduce(ctx, comm, sc + 4, 1));
        GKOB_CALL(vabi<V>::fused_init(ctx, n, r, z, p_ext, q, dinv, sc, ctl, work_.get_data(),
                                      max_iters_, res_kind_, iter_first_ ? 1 : 0, baseline_,
                                      (V)reduction_, 0));
        GKOB_CALL(cabi<V>::allreduce(ctx, comm, sc + 6, 2));
        GKOB_CALL(vabi<V>::fused_finish(ctx, sc, ctl, 1, baseline_ == 0 ? 3 : baseline_,
                                        (V)reduction_));
        auto enqueue_iteration = [&]() {
            GKOB_CALL(vabi<V>::fused_step_p(ctx, n, p_ext, z, sc, ctl));
            GKOB_CALL(cabi<V>::halo_exchange(ctx, comm, A_->get_halo(), p_ext, ctl));
            GKOB_CALL((viabi<V, I>::csr_spmv_dot(ctx, Al->get_plan(), n, n + ng, nnz,
                                                 Al->get_const_row_ptrs(), Al->get_const_col_idxs(),
                                                 Al->get_const_values(), p_ext, q, sc + 2,
                                                 work_.get_data(), ctl)));
            GKOB_CALL(cabi<V>::allreduce(ctx, comm, sc + 2, 1));
            GKOB_CALL(vabi<V>::fused_step_xr(ctx, n, xv, r, p_ext, q, z, dinv, sc, ctl,
                                             work_.get_data(), 0));
            GKOB_CALL(cabi<V>::allreduce(ctx, comm, sc + 6, 2));
            GKOB_CALL(vabi<V>::fused_finish(ctx, sc, ctl, 0, 0, V(0)));
        };
        if (!graph_ || graph_x_ != xv) {
            b200_graph_destroy(graph_);
            graph_ = nullptr;
            (void)Al->get_plan();
            GKOB_CALL(b200_graph_begin_capture(ctx));
            for (int k = 0; k < check_every_; ++k) enqueue_iteration();
            GKOB_CALL(b200_graph_end_capture(ctx, &graph_));
            graph_x_ = xv;
        }
        int32 h[8] = {0};
        exec_->copy_to_host(h, ctl_.get_const_data(), 8);
        while (h[0] == 0) {
            const int32 before = h[1];
            GKOB_CALL(b200_graph_launch(ctx, graph_));
            exec_->copy_to_host(h, ctl_.get_const_data(), 8);
            if (h[0] == 0 && h[1] == before)
                throw Error("fused CG: the device iteration made no progress");
        }
        num_iterations_ = h[1];
        status_ = (uint8)h[0];
        A_->get_communicator()->check();
    }
    int64 get_num_iterations() const { return num_iterations_; }
    uint8 get_stop_status() const { return status_; }

private:
    std::shared_ptr<const Executor> exec_;
    std::shared_ptr<const Matrix<V, I>> A_;
    int64 max_iters_;
    int res_kind_, baseline_;
    double reduction_;
    bool iter_first_;
    int check_every_;
    array<V> inv_diag_, sc_, work_;
    array<int32> ctl_;
    std::unique_ptr<Dense> ws_, one_, neg_one_;
    b200_graph* graph_ = nullptr;
    const V* graph_x_ = nullptr;
    int64 num_iterations_ = 0;
    uint8 status_ = 0;
};

}  // namespace distributed
}  // namespace gko_b200
