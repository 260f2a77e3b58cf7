// gko_b200_staging.hpp -- apply with HOST-resident vectors.
// The reference's LinOp::apply accepts operands that live on another executor and clones
// them onto the operator's executor for the call (include/ginkgo/core/base/lin_op.hpp:129-215,
// make_temporary_clone / make_temporary_output_clone): upload b, apply, download x, all in
// sequence.  staged_apply does the same job as a pipeline: two staging-buffer pairs on the
// device and the C ABI's b200_pipe_* (two copy streams + events), so the upload of call k+1,
// the kernels of call k and the download of call k-1 overlap.  apply() returns as soon as
// the work is enqueued; wait() makes the results visible on the host.  Pinned host buffers
// overlap fully, pageable ones are still correct.
#pragma once

namespace gko_b200 {

template <typename V>
class staged_apply {
public:
    using Dense = matrix::Dense<V>;
    staged_apply(std::shared_ptr<const LinOp> op, size_type num_rhs = 1)
        : op_(std::move(op)), exec_(op_->get_executor()), nrhs_(num_rhs)
    {
        const auto sz = op_->get_size();
        for (int s = 0; s < b200_pipe_num_slots(); ++s) {
            b_.push_back(Dense::create(exec_, dim2{sz.cols, nrhs_}));
            x_.push_back(Dense::create(exec_, dim2{sz.rows, nrhs_}));
        }
        // created AFTER the staging buffers: b200_pipe_create orders both copy streams behind
        // everything enqueued on the executor's stream so far, i.e. behind whatever used the
        // pool blocks these buffers were carved from
        GKOB_CALL(b200_pipe_create(exec_->ctx(), &pipe_));
    }
    // the destructor body runs before the members die: the copy streams are drained
    // (b200_pipe_destroy synchronises them) before b_ / x_ are returned to the pool
    ~staged_apply() { b200_pipe_destroy(pipe_); }
    staged_apply(const staged_apply&) = delete;
    staged_apply& operator=(const staged_apply&) = delete;

    // x_host = op(b_host); row-major cols x nrhs / rows x nrhs host arrays
    void apply(const V* b_host, V* x_host)
    {
        const int s = next_++ % b200_pipe_num_slots();
        const auto sz = op_->get_size();
        GKOB_CALL(b200_pipe_upload(pipe_, s, b_[s]->get_values(), b_host,
                                   sizeof(V) * sz.cols * nrhs_));
        GKOB_CALL(b200_pipe_begin_compute(pipe_, s));
        op_->apply(b_[s].get(), x_[s].get());
        GKOB_CALL(b200_pipe_end_compute(pipe_, s));
        GKOB_CALL(b200_pipe_download(pipe_, s, x_host, x_[s]->get_const_values(),
                                     sizeof(V) * sz.rows * nrhs_));
    }
    // the executor's stream waits for all outstanding transfers (so an event recorded on it,
    // or exec->synchronize(), covers the whole pipeline)
    void join() { GKOB_CALL(b200_pipe_join(pipe_)); }
    void wait()
    {
        join();
        exec_->synchronize();
    }

private:
    std::shared_ptr<const LinOp> op_;
    std::shared_ptr<const Executor> exec_;
    size_type nrhs_;
    b200_pipe* pipe_ = nullptr;
    std::vector<std::unique_ptr<Dense>> b_, x_;
    unsigned next_ = 0;
};

}  // namespace gko_b200
