// gko_b200_convert.hpp -- Csr::convert_to(Ell | Sellp | Coo | Hybrid) and
// Csr::sort_by_column_index, all on the device.  Same sequence of kernel
// calls as the reference host code (core/matrix/csr.cpp:285-300 Coo, :419-441 Hybrid,
// :452-472 Sellp, :506-530 Ell, :1402-1406 sort); the kernels are the C-ABI entry points of
// include/ginkgo_b200.h ("CSR -> ELL / SELL-P / Hybrid").  Included by gko_b200.hpp.
#pragma once

namespace gko_b200 {
namespace matrix {

template <typename V, typename I>
void Csr<V, I>::convert_to(Ell<V, I>* result) const
{
    auto ctx = exec_->ctx();
    int64 max_nnz = 0;
    GKOB_CALL((viabi<V, I>::ell_compute_max_row_nnz(ctx, get_const_row_ptrs(), size_.rows, &max_nnz)));
    const size_type width = (size_type)max_nnz, stride = size_.rows;
    array<V> vals(exec_, width * stride);
    array<I> cols(exec_, width * stride);
    GKOB_CALL((viabi<V, I>::csr_convert_to_ell(ctx, size_.rows, get_const_row_ptrs(),
                                               get_const_col_idxs(), get_const_values(), width,
                                               stride, cols.get_data(), vals.get_data())));
    result->size_ = size_;
    result->width_ = width;
    result->stride_ = stride;
    result->values_ = std::move(vals);
    result->col_idxs_ = std::move(cols);
}

template <typename V, typename I>
void Csr<V, I>::convert_to(Sellp<V, I>* result) const
{
    auto ctx = exec_->ctx();
    const size_type slice_size = result->get_slice_size();
    const size_type stride_factor = result->get_stride_factor();
    const size_type num_slices = (size_.rows + slice_size - 1) / slice_size;
    array<std::uint64_t> sets(exec_, num_slices + 1), lens(exec_, num_slices);
    GKOB_CALL((viabi<V, I>::sellp_compute_slice_sets(ctx, get_const_row_ptrs(), size_.rows,
                                                     slice_size, stride_factor, sets.get_data(),
                                                     lens.get_data())));
    std::uint64_t total_cols = 0;  // the reference's copy_val_to_host(slice_sets + num_slices)
    exec_->copy_to_host(&total_cols, sets.get_const_data() + num_slices, 1);
    array<V> vals(exec_, total_cols * slice_size);
    array<I> cols(exec_, total_cols * slice_size);
    GKOB_CALL((viabi<V, I>::csr_convert_to_sellp(ctx, size_.rows, slice_size, sets.get_const_data(),
                                                 lens.get_const_data(), get_const_row_ptrs(),
                                                 get_const_col_idxs(), get_const_values(),
                                                 cols.get_data(), vals.get_data())));
    result->size_ = size_;
    result->sets_ = std::move(sets);
    result->lens_ = std::move(lens);
    result->values_ = std::move(vals);
    result->col_idxs_ = std::move(cols);
}

template <typename V, typename I>
void Csr<V, I>::convert_to(Coo<V, I>* result) const
{
    auto ctx = exec_->ctx();
    const size_type nnz = get_num_stored_elements();
    array<V> vals(exec_, nnz);
    array<I> cols(exec_, nnz), rows(exec_, nnz);
    exec_->copy(vals.get_data(), get_const_values(), nnz);
    exec_->copy(cols.get_data(), get_const_col_idxs(), nnz);
    GKOB_CALL((viabi<V, I>::convert_ptrs_to_idxs(ctx, get_const_row_ptrs(), size_.rows, rows.get_data())));
    b200_coo_plan_destroy(result->plan_);
    result->plan_ = nullptr;
    result->size_ = size_;
    result->values_ = std::move(vals);
    result->col_idxs_ = std::move(cols);
    result->row_idxs_ = std::move(rows);
}

template <typename V, typename I>
void Csr<V, I>::convert_to(Hybrid<V, I>* result) const
{
    auto ctx = exec_->ctx();
    const size_type n = size_.rows;
    size_type ell_lim = result->get_strategy().compute_ell_num_stored_elements_per_row(
        exec_.get(), get_const_row_ptrs(), n);
    if (ell_lim > size_.cols) ell_lim = size_.cols;  // core/matrix/csr.cpp:428-431
    array<int64> coo_row_ptrs(exec_, n + 1);
    GKOB_CALL((viabi<V, I>::csr_compute_hybrid_coo_row_ptrs(ctx, get_const_row_ptrs(), n, ell_lim,
                                                            coo_row_ptrs.get_data())));
    int64 coo_nnz = 0;
    exec_->copy_to_host(&coo_nnz, coo_row_ptrs.get_const_data() + n, 1);
    const size_type stride = n;
    array<V> evals(exec_, ell_lim * stride), cvals(exec_, coo_nnz);
    array<I> ecols(exec_, ell_lim * stride), ccols(exec_, coo_nnz), crows(exec_, coo_nnz);
    GKOB_CALL((viabi<V, I>::csr_convert_to_hybrid(
        ctx, n, get_const_row_ptrs(), get_const_col_idxs(), get_const_values(), ell_lim, stride,
        ecols.get_data(), evals.get_data(), coo_row_ptrs.get_const_data(), crows.get_data(),
        ccols.get_data(), cvals.get_data())));
    auto ell = result->ell_.get();
    ell->size_ = size_;
    ell->width_ = ell_lim;
    ell->stride_ = stride;
    ell->values_ = std::move(evals);
    ell->col_idxs_ = std::move(ecols);
    auto coo = result->coo_.get();
    b200_coo_plan_destroy(coo->plan_);
    coo->plan_ = nullptr;
    coo->size_ = size_;
    coo->values_ = std::move(cvals);
    coo->col_idxs_ = std::move(ccols);
    coo->row_idxs_ = std::move(crows);
    result->size_ = size_;
}

template <typename V, typename I>
void Csr<V, I>::sort_by_column_index()
{
    GKOB_CALL((viabi<V, I>::csr_sort_by_column_index(exec_->ctx(), size_.rows, get_const_row_ptrs(),
                                                     col_idxs_.get_data(), values_.get_data())));
}

// ---- ReadableFromMatrixData / WritableToMatrixData of the other formats ---------------------
// read: through Csr (Csr::read, then the device conversion) -- the same layout the reference's
// own read produces (core/matrix/{ell,sellp,coo,hybrid}.cpp `read`: widest row / slice sets /
// strategy split computed from the row lengths).  write: the reference's loops
// (core/matrix/ell.cpp `write`, sellp.cpp, coo.cpp, hybrid.cpp) over host copies of the arrays.
template <typename V, typename I>
void Ell<V, I>::read(const matrix_data<V, I>& data)
{
    auto tmp = Csr<V, I>::create(exec_);
    tmp->read(data);
    tmp->convert_to(this);
}
template <typename V, typename I>
void Sellp<V, I>::read(const matrix_data<V, I>& data)
{
    auto tmp = Csr<V, I>::create(exec_);
    tmp->read(data);
    tmp->convert_to(this);
}
template <typename V, typename I>
void Coo<V, I>::read(const matrix_data<V, I>& data)
{
    auto tmp = Csr<V, I>::create(exec_);
    tmp->read(data);
    tmp->convert_to(this);
}
template <typename V, typename I>
void Hybrid<V, I>::read(const matrix_data<V, I>& data)
{
    auto tmp = Csr<V, I>::create(exec_);
    tmp->read(data);
    tmp->convert_to(this);
}

template <typename V, typename I>
void Ell<V, I>::write(matrix_data<V, I>& data) const
{
    const auto vals = values_.to_host();
    const auto cols = col_idxs_.to_host();
    data = matrix_data<V, I>(size_);
    for (size_type row = 0; row < size_.rows; ++row)
        for (size_type i = 0; i < width_; ++i) {
            const I col = cols[row + i * stride_];
            if (col != I(-1)) data.nonzeros.push_back({(I)row, col, vals[row + i * stride_]});
        }
}
template <typename V, typename I>
void Sellp<V, I>::write(matrix_data<V, I>& data) const
{
    const auto vals = values_.to_host();
    const auto cols = col_idxs_.to_host();
    const auto sets = sets_.to_host();
    const auto lens = lens_.to_host();
    data = matrix_data<V, I>(size_);
    const size_type num_slices = lens.size();
    for (size_type slice = 0; slice < num_slices; ++slice)
        for (size_type r = 0; r < slice_size_; ++r) {
            const size_type row = slice * slice_size_ + r;
            if (row >= size_.rows) break;
            for (size_type i = 0; i < lens[slice]; ++i) {
                const size_type at = (sets[slice] + i) * slice_size_ + r;
                if (cols[at] != I(-1)) data.nonzeros.push_back({(I)row, cols[at], vals[at]});
            }
        }
}
template <typename V, typename I>
void Coo<V, I>::write(matrix_data<V, I>& data) const
{
    const auto vals = values_.to_host();
    const auto cols = col_idxs_.to_host();
    const auto rows = row_idxs_.to_host();
    data = matrix_data<V, I>(size_);
    for (size_type k = 0; k < vals.size(); ++k) data.nonzeros.push_back({rows[k], cols[k], vals[k]});
}
template <typename V, typename I>
void Hybrid<V, I>::write(matrix_data<V, I>& data) const
{
    const auto evals = ell_->values_.to_host();
    const auto ecols = ell_->col_idxs_.to_host();
    const auto cvals = coo_->values_.to_host();
    const auto ccols = coo_->col_idxs_.to_host();
    const auto crows = coo_->row_idxs_.to_host();
    const size_type width = ell_->width_, stride = ell_->stride_;
    data = matrix_data<V, I>(size_);
    size_type coo_at = 0;
    for (size_type row = 0; row < size_.rows; ++row) {
        for (size_type i = 0; i < width; ++i) {
            const I col = ecols[row + i * stride];
            if (col != I(-1)) data.nonzeros.push_back({(I)row, col, evals[row + i * stride]});
        }
        while (coo_at < cvals.size() && (size_type)crows[coo_at] == row) {
            data.nonzeros.push_back({(I)row, ccols[coo_at], cvals[coo_at]});
            ++coo_at;
        }
    }
}

}  // namespace matrix
}  // namespace gko_b200
