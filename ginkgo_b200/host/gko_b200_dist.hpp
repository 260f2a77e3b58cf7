// gko_b200_dist.hpp -- multi-GPU host layer: the NCCL / peer-memory counterpart of the reference's
// experimental::distributed classes on this path, one process per GPU:
//   communicator                      experimental::mpi::communicator (include/ginkgo/core/base/mpi.hpp)
//   Partition<L, G>, index_map<L, G>  include/ginkgo/core/distributed/{partition,index_map}.hpp
//   Vector<V>                         include/ginkgo/core/distributed/vector.hpp (core/distributed/vector.cpp:480-534)
//   Matrix<V, I>                      core/distributed/matrix.cpp:300-380 (read_distributed), :450-509 (apply);
//                                     a LinOp over the local rows, so every solver of gko_b200_solvers.hpp runs on it
//   preconditioner::Schwarz<V, I>     core/distributed/preconditioner/schwarz.cpp (one level)
//   Cg<V, I>                          the fused device-resident CG iteration with the exchange and the
//                                     all-reduces inside the CUDA graph (core/solver/cg.cpp driven through
//                                     precision_dispatch_real_complex_distributed)
// Rank p owns the rows of part p of the row partition; its local block has the columns numbered
// [owned columns | ghost columns ordered by (owner, global index)].
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "gko_b200.hpp"
#include "gko_b200_io.hpp"

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
    static constexpr auto halo_exchange_inplace = b200_halo_exchange_inplace_f64;
    static constexpr auto halo_exchange_staged_begin = b200_halo_exchange_staged_begin_f64;
};
template <>
struct cabi<float> {
    static constexpr auto allreduce = b200_comm_allreduce_sum_f32;
    static constexpr auto halo_exchange = b200_halo_exchange_f32;
    static constexpr auto halo_exchange_inplace = b200_halo_exchange_inplace_f32;
    static constexpr auto halo_exchange_staged_begin = b200_halo_exchange_staged_begin_f32;
};

// the all-reduce behind a distributed vector's dots and norms
template <typename V>
class comm_reducer : public matrix::reducer<V> {
public:
    comm_reducer(std::shared_ptr<const Executor> exec, std::shared_ptr<communicator> comm)
        : exec_(std::move(exec)), comm_(std::move(comm))
    {}
    void sum(V* device_values, size_type count) const override
    {
        if (comm_->size() > 1 && count)
            GKOB_CALL(cabi<V>::allreduce(exec_->ctx(), comm_->get(), device_values, (int64)count));
    }
    const std::shared_ptr<communicator>& get_communicator() const { return comm_; }

private:
    std::shared_ptr<const Executor> exec_;
    std::shared_ptr<communicator> comm_;
};

// experimental::distributed::Vector (include/ginkgo/core/distributed/vector.hpp): the rows a rank
// owns of a global vector.  It IS the Dense of its local rows (get_local_vector() == this), so
// every solver takes it as is; compute_dot / compute_conj_dot / compute_norm2 /
// compute_squared_norm2 add the sum over the ranks (core/distributed/vector.cpp:480-534), and
// the vectors a solver creates next to it reduce the same way (Dense::create_like).
// A solver on a distributed::Matrix must be given distributed::Vector operands: with plain Dense
// operands its dots and norms would stay local to each rank.
template <typename V>
class Vector : public matrix::Dense<V> {
    using Dense = matrix::Dense<V>;

public:
    static std::unique_ptr<Vector> create(std::shared_ptr<const Executor> exec, std::shared_ptr<communicator> comm,
                                          dim2 global_size, dim2 local_size)
    {
        if (global_size.cols != local_size.cols)
            throw DimensionMismatch("distributed::Vector: global and local column counts differ");
        return std::unique_ptr<Vector>(new Vector(exec, comm, global_size, local_size,
                                                  array<V>(exec, local_size.rows * local_size.cols),
                                                  local_size.cols));
    }
    // non-owning view of the local rows in device memory
    static std::unique_ptr<Vector> create_view(std::shared_ptr<const Executor> exec,
                                               std::shared_ptr<communicator> comm, dim2 global_size,
                                               dim2 local_size, V* device_ptr, size_type stride)
    {
        return std::unique_ptr<Vector>(new Vector(exec, comm, global_size, local_size,
                                                  array<V>::view(exec, local_size.rows * stride, device_ptr),
                                                  stride));
    }
    dim2 get_global_size() const { return global_size_; }
    const Dense* get_local_vector() const { return this; }
    Dense* get_local_vector() { return this; }
    std::shared_ptr<communicator> get_communicator() const { return comm_; }

private:
    Vector(std::shared_ptr<const Executor> exec, std::shared_ptr<communicator> comm, dim2 global_size,
           dim2 local_size, array<V> values, size_type stride)
        : Dense(exec, local_size, std::move(values), stride), global_size_(global_size), comm_(comm)
    {
        this->set_reducer(std::make_shared<comm_reducer<V>>(exec, comm));
    }
    dim2 global_size_;
    std::shared_ptr<communicator> comm_;
};

// C-ABI tables of the distributed set-up kernels, by global / (local, global) / (value, local,
// global) types
template <typename I>
struct iabi;
template <>
struct iabi<int32> {
    static constexpr auto convert_idxs_to_ptrs = b200_convert_idxs_to_ptrs_i32;
};
template <>
struct iabi<int64> {
    static constexpr auto convert_idxs_to_ptrs = b200_convert_idxs_to_ptrs_i64;
};
template <typename G>
struct gabi;
template <typename L, typename G>
struct lgabi;
template <typename V, typename L, typename G>
struct vlgabi;
#define GKOB_G(G, S)                                                                                  \
    template <>                                                                                       \
    struct gabi<G> {                                                                                  \
        static constexpr auto build_ranges_from_global_size =                                         \
            b200_partition_build_ranges_from_global_size_##S;                                         \
        static constexpr auto build_from_contiguous = b200_partition_build_from_contiguous_##S;       \
        static constexpr auto build_from_mapping = b200_partition_build_from_mapping_##S;             \
        static constexpr auto classify_entries = b200_dist_classify_entries_##S;                      \
        static constexpr auto index_map_mark = b200_index_map_mark_##S;                               \
        static constexpr auto index_map_rank = b200_index_map_rank_##S;                               \
    };
GKOB_G(int32, i32)
GKOB_G(int64, i64)
#define GKOB_LG(L, LS, G, GS)                                                                         \
    template <>                                                                                       \
    struct lgabi<L, G> {                                                                              \
        static constexpr auto build_starting_indices = b200_partition_build_starting_indices_##LS##_##GS; \
        static constexpr auto index_map_fill = b200_index_map_fill_##LS##_##GS;                       \
        static constexpr auto index_map_map_to_local = b200_index_map_map_to_local_##LS##_##GS;       \
    };
GKOB_LG(int32, i32, int32, i32)
GKOB_LG(int32, i32, int64, i64)
GKOB_LG(int64, i64, int64, i64)
#define GKOB_VLG(V, VS, L, LS, G, GS)                                                                 \
    template <>                                                                                       \
    struct vlgabi<V, L, G> {                                                                          \
        static constexpr auto separate_fill = b200_dist_separate_fill_##VS##_##LS##_##GS;             \
        static constexpr auto kept_fill = b200_dist_kept_fill_##VS##_##LS##_##GS;                     \
        static constexpr auto vector_build_local = b200_dist_vector_build_local_##VS##_##LS##_##GS;   \
    };
GKOB_VLG(double, f64, int32, i32, int32, i32)
GKOB_VLG(double, f64, int32, i32, int64, i64)
GKOB_VLG(double, f64, int64, i64, int64, i64)
GKOB_VLG(float, f32, int32, i32, int32, i32)
GKOB_VLG(float, f32, int32, i32, int64, i64)
GKOB_VLG(float, f32, int64, i64, int64, i64)

using comm_index_type = int32;
enum class index_space { local = 0, non_local = 1, combined = 2 };

// experimental::distributed::Partition (include/ginkgo/core/distributed/partition.hpp:83-290,
// core/distributed/partition.cpp): ranges of global indices, each owned by one part; all
// arrays live on the device and are built by device kernels.
template <typename L = int32, typename G = int64>
class Partition {
public:
    using local_index_type = L;
    using global_index_type = G;

    // one range per run of equal owners in `mapping` (device array of part ids)
    static std::unique_ptr<Partition> build_from_mapping(std::shared_ptr<const Executor> exec,
                                                         const array<comm_index_type>& mapping,
                                                         comm_index_type num_parts)
    {
        int64 num_ranges = 0;
        GKOB_CALL(b200_partition_count_ranges(exec->ctx(), (int64)mapping.get_size(),
                                              mapping.get_const_data(), &num_ranges));
        auto p = std::unique_ptr<Partition>(new Partition(exec, num_parts, (size_type)num_ranges));
        GKOB_CALL(gabi<G>::build_from_mapping(exec->ctx(), (int64)mapping.get_size(),
                                              mapping.get_const_data(), p->offsets_.get_data(),
                                              p->part_ids_.get_data()));
        p->finalize_construction();
        return p;
    }
    // ranges[i], ranges[i+1] bound range i; part_ids (optional) names its owner, default part i
    static std::unique_ptr<Partition> build_from_contiguous(std::shared_ptr<const Executor> exec,
                                                            const array<G>& ranges,
                                                            const array<comm_index_type>& part_ids = {})
    {
        if (ranges.get_size() == 0) throw BadDimension("Partition: ranges needs at least one entry");
        const size_type n = ranges.get_size() - 1;
        if (part_ids.get_size() != 0 && part_ids.get_size() != n)
            throw BadDimension("Partition: part_ids must have one entry per range");
        auto p = std::unique_ptr<Partition>(new Partition(exec, (comm_index_type)n, n));
        GKOB_CALL(gabi<G>::build_from_contiguous(exec->ctx(), (int64)n, ranges.get_const_data(),
                                                 part_ids.get_size() ? part_ids.get_const_data() : nullptr,
                                                 p->offsets_.get_data(), p->part_ids_.get_data()));
        p->finalize_construction();
        return p;
    }
    static std::unique_ptr<Partition> build_from_global_size_uniform(std::shared_ptr<const Executor> exec,
                                                                     comm_index_type num_parts,
                                                                     G global_size)
    {
        array<G> ranges(exec, (size_type)num_parts + 1);
        // zero parts: the single bound is 0 whatever the size (partition.cpp:106-111)
        GKOB_CALL(gabi<G>::build_ranges_from_global_size(exec->ctx(), num_parts,
                                                         num_parts ? (int64)global_size : 0,
                                                         ranges.get_data()));
        return build_from_contiguous(exec, ranges);
    }

    size_type get_size() const { return size_; }
    size_type get_num_ranges() const noexcept { return offsets_.get_size() - 1; }
    comm_index_type get_num_parts() const noexcept { return num_parts_; }
    comm_index_type get_num_empty_parts() const noexcept { return num_empty_parts_; }
    // device pointers
    const G* get_range_bounds() const noexcept { return offsets_.get_const_data(); }
    const comm_index_type* get_part_ids() const noexcept { return part_ids_.get_const_data(); }
    const L* get_range_starting_indices() const noexcept { return starting_indices_.get_const_data(); }
    const L* get_part_sizes() const noexcept { return part_sizes_.get_const_data(); }
    L get_part_size(comm_index_type part) const
    {
        if (part < 0 || part >= num_parts_) throw OutOfBounds("Partition::get_part_size");
        return host_part_sizes_[part];
    }
    bool has_connected_parts() const
    {
        return (size_type)(num_parts_ - num_empty_parts_) == get_num_ranges();
    }
    bool has_ordered_parts() const
    {
        if (!has_connected_parts()) return false;
        int32 res = 0;
        GKOB_CALL(b200_partition_has_ordered_parts(exec_->ctx(), (int64)get_num_ranges(),
                                                   part_ids_.get_const_data(), &res));
        return res != 0;
    }
    std::shared_ptr<const Executor> get_executor() const { return exec_; }

private:
    Partition(std::shared_ptr<const Executor> exec, comm_index_type num_parts, size_type num_ranges)
        : exec_(exec),
          num_parts_(num_parts),
          offsets_(exec, num_ranges + 1),
          starting_indices_(exec, num_ranges),
          part_sizes_(exec, (size_type)num_parts),
          part_ids_(exec, num_ranges)
    {}
    void finalize_construction()
    {
        int32 empty = 0;
        GKOB_CALL((lgabi<L, G>::build_starting_indices(
            exec_->ctx(), (int64)get_num_ranges(), num_parts_, offsets_.get_const_data(),
            part_ids_.get_const_data(), starting_indices_.get_data(), part_sizes_.get_data(), &empty)));
        num_empty_parts_ = empty;
        G last = 0;
        exec_->copy_to_host(&last, offsets_.get_const_data() + get_num_ranges(), 1);
        size_ = (size_type)last;
        host_part_sizes_ = part_sizes_.to_host();
    }
    std::shared_ptr<const Executor> exec_;
    comm_index_type num_parts_;
    comm_index_type num_empty_parts_ = 0;
    size_type size_ = 0;
    array<G> offsets_;
    array<L> starting_indices_;
    array<L> part_sizes_;
    array<comm_index_type> part_ids_;
    std::vector<L> host_part_sizes_;
};

// experimental::distributed::index_map (include/ginkgo/core/distributed/index_map.hpp:60-200):
// the remote indices a rank touches, ordered by (owning part, global index), and the maps
// between the global and the local / non-local / combined index spaces.
template <typename L = int32, typename G = int64>
class index_map {
public:
    using partition_type = Partition<L, G>;
    // `connections`: device array of global indices; entries owned by `rank` are ignored
    index_map(std::shared_ptr<const Executor> exec, std::shared_ptr<const partition_type> part,
              comm_index_type rank, const array<G>& connections)
        : exec_(exec), part_(part), rank_(rank)
    {
        if (rank < 0 || rank >= part->get_num_parts()) throw OutOfBounds("index_map: rank");
        const int64 gs = (int64)part->get_size(), nr = (int64)part->get_num_ranges();
        const int64 words = (gs + 31) / 32;
        bitmap_ = array<uint32>(exec, (size_type)words + 1);
        word_rank_ = array<int64>(exec, (size_type)words + 1);
        range_offsets_ = array<int64>(exec, (size_type)nr);
        array<int64> sizes(exec, (size_type)part->get_num_parts());
        auto ctx = exec->ctx();
        GKOB_CALL(gabi<G>::index_map_mark(ctx, gs, nr, part->get_range_bounds(), part->get_part_ids(), rank,
                                          (int64)connections.get_size(), connections.get_const_data(),
                                          bitmap_.get_data()));
        int64 num_remote = 0;
        GKOB_CALL(gabi<G>::index_map_rank(ctx, gs, nr, part->get_num_parts(), part->get_range_bounds(),
                                          part->get_part_ids(), bitmap_.get_const_data(),
                                          word_rank_.get_data(), range_offsets_.get_data(), sizes.get_data(),
                                          &num_remote));
        remote_sizes_ = sizes.to_host();
        remote_global_idxs_ = array<G>(exec, (size_type)num_remote);
        remote_local_idxs_ = array<L>(exec, (size_type)num_remote);
        GKOB_CALL((lgabi<L, G>::index_map_fill(
            ctx, gs, nr, part->get_range_bounds(), part->get_part_ids(), part->get_range_starting_indices(),
            bitmap_.get_const_data(), word_rank_.get_const_data(), range_offsets_.get_const_data(),
            remote_global_idxs_.get_data(), remote_local_idxs_.get_data(), nullptr)));
    }
    size_type get_global_size() const { return part_->get_size(); }
    size_type get_local_size() const { return (size_type)part_->get_part_size(rank_); }
    size_type get_non_local_size() const { return remote_global_idxs_.get_size(); }
    // flat device arrays ordered by (part, global index); segment p has get_remote_sizes()[p] entries
    const array<G>& get_remote_global_idxs() const { return remote_global_idxs_; }
    const array<L>& get_remote_local_idxs() const { return remote_local_idxs_; }
    const std::vector<int64>& get_remote_sizes() const { return remote_sizes_; }
    // the parts this rank receives from (the reference's get_remote_target_ids)
    std::vector<comm_index_type> get_remote_target_ids() const
    {
        std::vector<comm_index_type> ids;
        for (size_type p = 0; p < remote_sizes_.size(); ++p)
            if (remote_sizes_[p]) ids.push_back((comm_index_type)p);
        return ids;
    }
    // global -> local / non-local / combined index; invalid_index (-1) where not representable
    array<L> map_to_local(const array<G>& global_ids, index_space is) const
    {
        array<L> out(exec_, global_ids.get_size());
        map_to_local(global_ids.get_const_data(), global_ids.get_size(), is, out.get_data());
        return out;
    }
    void map_to_local(const G* global_ids, size_type m, index_space is, L* local_ids) const
    {
        GKOB_CALL((lgabi<L, G>::index_map_map_to_local(
            exec_->ctx(), (int64)part_->get_size(), (int64)part_->get_num_ranges(), part_->get_range_bounds(),
            part_->get_part_ids(), part_->get_range_starting_indices(), bitmap_.get_const_data(),
            word_rank_.get_const_data(), range_offsets_.get_const_data(), rank_,
            (L)part_->get_part_size(rank_), (int32)is, (int64)m, global_ids, local_ids)));
    }

private:
    std::shared_ptr<const Executor> exec_;
    std::shared_ptr<const partition_type> part_;
    comm_index_type rank_;
    array<uint32> bitmap_;
    array<int64> word_rank_, range_offsets_;
    array<G> remote_global_idxs_;
    array<L> remote_local_idxs_;
    std::vector<int64> remote_sizes_;
};

// What one rank owns of a global matrix, in the combined index space of its columns
// [local columns | remote columns ordered by (part, global index)]: the communication-free
// part of Matrix::read_distributed (core/distributed/matrix.cpp:343-373).  `data` must be
// sorted row-major like the reference requires of device_matrix_data handed to Csr::read.
template <typename V, typename I, typename G>
struct local_assembly {
    std::unique_ptr<matrix::Csr<V, I>> local;  // n_local_rows x (n_local_cols + n_ghost)
    size_type n_local_rows = 0, n_local_cols = 0, n_ghost = 0;
    std::unique_ptr<index_map<I, G>> imap;     // remote columns of this rank
    // optional: the square block of the owned columns alone (the reference's local_mtx), for
    // local solvers (Schwarz)
    std::shared_ptr<matrix::Csr<V, I>> local_only;
};

template <typename V, typename I, typename G>
local_assembly<V, I, G> assemble_local(std::shared_ptr<const Executor> exec, const matrix_data<V, G>& data,
                                       std::shared_ptr<const Partition<I, G>> row_part,
                                       std::shared_ptr<const Partition<I, G>> col_part,
                                       comm_index_type rank, bool keep_local_block = false)
{
    if (data.size.rows != row_part->get_size() || data.size.cols != col_part->get_size())
        throw DimensionMismatch("read_distributed: the partitions must cover the matrix");
    if (row_part->get_num_parts() != col_part->get_num_parts())
        throw DimensionMismatch("read_distributed: row and column partition need the same parts");
    if (rank < 0 || rank >= row_part->get_num_parts()) throw OutOfBounds("read_distributed: rank");
    auto ctx = exec->ctx();
    const size_type nnz = data.nonzeros.size();
    std::vector<G> hr(nnz), hc(nnz);
    std::vector<V> hv(nnz);
    for (size_type i = 0; i < nnz; ++i) {
        hr[i] = data.nonzeros[i].row;
        hc[i] = data.nonzeros[i].column;
        hv[i] = data.nonzeros[i].value;
    }
    array<G> rows(exec, hr), cols(exec, hc);
    array<V> vals(exec, hv);
    array<uint8> cls(exec, nnz);
    array<int64> lrank(exec, nnz + 1), nrank(exec, nnz + 1);
    int64 n_loc = 0, n_non = 0;
    GKOB_CALL(gabi<G>::classify_entries(
        ctx, (int64)nnz, rows.get_const_data(), cols.get_const_data(), (int64)row_part->get_num_ranges(),
        row_part->get_range_bounds(), row_part->get_part_ids(), (int64)col_part->get_num_ranges(),
        col_part->get_range_bounds(), col_part->get_part_ids(), rank, cls.get_data(), lrank.get_data(),
        nrank.get_data(), &n_loc, &n_non));
    const size_type kept = (size_type)(n_loc + n_non);
    array<I> krows(exec, kept);
    array<G> kcols(exec, kept);
    array<V> kvals(exec, kept);
    GKOB_CALL((vlgabi<V, I, G>::kept_fill(
        ctx, (int64)nnz, rows.get_const_data(), cols.get_const_data(), vals.get_const_data(),
        (int64)row_part->get_num_ranges(), row_part->get_range_bounds(), row_part->get_range_starting_indices(),
        cls.get_const_data(), lrank.get_const_data(), nrank.get_const_data(), krows.get_data(),
        kcols.get_data(), kvals.get_data())));
    local_assembly<V, I, G> out;
    out.imap.reset(new index_map<I, G>(exec, col_part, rank, kcols));
    out.n_local_rows = (size_type)row_part->get_part_size(rank);
    out.n_local_cols = (size_type)col_part->get_part_size(rank);
    out.n_ghost = out.imap->get_non_local_size();
    array<I> lcols(exec, kept);
    out.imap->map_to_local(kcols.get_const_data(), kept, index_space::combined, lcols.get_data());
    array<I> row_ptrs(exec, out.n_local_rows + 1);
    GKOB_CALL(iabi<I>::convert_idxs_to_ptrs(ctx, krows.get_const_data(), (int64)kept, (int64)out.n_local_rows,
                                            row_ptrs.get_data()));
    out.local = matrix::Csr<V, I>::create(exec, dim2{out.n_local_rows, out.n_local_cols + out.n_ghost},
                                          std::move(kvals), std::move(lcols), std::move(row_ptrs));
    if (keep_local_block) {
        // separate_local_nonlocal's local arrays (the non-local ones are not needed here)
        array<I> lr(exec, (size_type)n_loc), lc(exec, (size_type)n_loc), nr(exec, (size_type)n_non);
        array<G> nc(exec, (size_type)n_non);
        array<V> lv(exec, (size_type)n_loc), nv(exec, (size_type)n_non);
        GKOB_CALL((vlgabi<V, I, G>::separate_fill(
            ctx, (int64)nnz, rows.get_const_data(), cols.get_const_data(), vals.get_const_data(),
            (int64)row_part->get_num_ranges(), row_part->get_range_bounds(),
            row_part->get_range_starting_indices(), (int64)col_part->get_num_ranges(),
            col_part->get_range_bounds(), col_part->get_range_starting_indices(), cls.get_const_data(),
            lrank.get_const_data(), nrank.get_const_data(), lr.get_data(), lc.get_data(), lv.get_data(),
            nr.get_data(), nc.get_data(), nv.get_data())));
        array<I> lptrs(exec, out.n_local_rows + 1);
        GKOB_CALL(iabi<I>::convert_idxs_to_ptrs(ctx, lr.get_const_data(), n_loc, (int64)out.n_local_rows,
                                                lptrs.get_data()));
        out.local_only = matrix::Csr<V, I>::create(exec, dim2{out.n_local_rows, out.n_local_cols},
                                                   std::move(lv), std::move(lc), std::move(lptrs));
    }
    return out;
}

// experimental::distributed::Vector::read_distributed (core/distributed/vector.cpp:250-275): the
// rows of `partition`'s part comm->rank() of a global multi-vector given as (unique) triplets
template <typename V, typename L, typename G>
std::unique_ptr<Vector<V>> read_distributed_vector(std::shared_ptr<const Executor> exec,
                                                   std::shared_ptr<communicator> comm,
                                                   const matrix_data<V, G>& data,
                                                   std::shared_ptr<const Partition<L, G>> partition)
{
    if (data.size.rows != partition->get_size())
        throw DimensionMismatch("read_distributed: the partition must cover the rows of the vector");
    if (partition->get_num_parts() != comm->size())
        throw DimensionMismatch("read_distributed: one part per rank of the communicator");
    const size_type nnz = data.nonzeros.size(), cols = data.size.cols;
    std::vector<G> hr(nnz), hc(nnz);
    std::vector<V> hv(nnz);
    for (size_type i = 0; i < nnz; ++i) {
        hr[i] = data.nonzeros[i].row;
        hc[i] = data.nonzeros[i].column;
        hv[i] = data.nonzeros[i].value;
    }
    array<G> rows(exec, hr), cs(exec, hc);
    array<V> vals(exec, hv);
    const size_type n_local = (size_type)partition->get_part_size(comm->rank());
    auto v = Vector<V>::create(exec, comm, data.size, dim2{n_local, cols});
    v->fill(V(0));
    GKOB_CALL((vlgabi<V, L, G>::vector_build_local(
        exec->ctx(), (int64)nnz, rows.get_const_data(), cs.get_const_data(), vals.get_const_data(),
        (int64)partition->get_num_ranges(), partition->get_range_bounds(), partition->get_part_ids(),
        partition->get_range_starting_indices(), comm->rank(), v->get_values(), (int64)v->get_stride())));
    return v;
}

// S[q * P + p] = entries rank q receives from rank p.  What `rank` sends: to peer q the
// S[q][rank] entries that start at offset sum_{p < rank} S[q][p] of q's remote list.
struct send_layout {
    std::vector<int64> send_counts, source_offsets;
};
inline send_layout compute_send_layout(const std::vector<int64>& S, int P, int rank)
{
    send_layout l;
    l.send_counts.resize(P);
    l.source_offsets.resize(P);
    for (int q = 0; q < P; ++q) {
        l.send_counts[q] = S[(size_t)q * P + rank];
        int64 off = 0;
        for (int p = 0; p < rank; ++p) off += S[(size_t)q * P + p];
        l.source_offsets[q] = off;
    }
    return l;
}

// distributed::Matrix: local rows, columns numbered into the extended vector
// [n_local owned | n_ghost received]; apply = halo exchange + one local SpMV.
template <typename V, typename I>
class Matrix : public LinOp {
public:
    Matrix(std::shared_ptr<const Executor> exec, std::shared_ptr<communicator> comm,
           std::unique_ptr<matrix::Csr<V, I>> local, size_type n_ghost,
           const std::vector<int64>& send_counts, const std::vector<int64>& recv_counts,
           const int32* send_idx_dev, int64 n_local_cols = -1)
        : LinOp(exec, dim2{local->get_size().rows,
                           n_local_cols < 0 ? local->get_size().rows : (size_type)n_local_cols}),
          comm_(comm), local_(std::move(local)), n_ghost_(n_ghost)
    {
        // square row/column partition unless told otherwise
        n_local_cols_ = size_.cols;
        if (local_->get_size().cols != n_local_cols_ + n_ghost)
            throw BadDimension("distributed::Matrix: local block must be n_local x (n_local_cols+n_ghost)");
        GKOB_CALL(b200_halo_create(exec->ctx(), comm->size(), (int64)n_local_cols_, n_ghost,
                                   send_counts.data(), recv_counts.data(), send_idx_dev,
                                   (int32)sizeof(V), &halo_));
        if (comm->use_p2p() &&
            b200_halo_enable_p2p(exec->ctx(), comm->get(), halo_) != B200_OK)
            std::fprintf(stderr, "[gko_b200] rank %d: peer-memory halo unavailable (%s), using NCCL\n",
                         comm->rank(), b200_last_error());
    }
    ~Matrix()
    {
        b200_csr_plan_destroy(owner_plan_);
        b200_halo_destroy(halo_);
    }

    // experimental::distributed::Matrix::read_distributed (core/distributed/matrix.cpp:300-380):
    // every rank passes the (row-major sorted) global matrix or at least its own rows; the rows
    // of `row_part`'s part comm->rank() are kept, their columns renumbered into the combined
    // index space, and the ranks exchange which of their entries the others need.  Collective.
    template <typename G>
    static std::shared_ptr<Matrix> read_distributed(std::shared_ptr<const Executor> exec,
                                                    std::shared_ptr<communicator> comm,
                                                    const matrix_data<V, G>& data,
                                                    std::shared_ptr<const Partition<I, G>> row_part,
                                                    std::shared_ptr<const Partition<I, G>> col_part = nullptr,
                                                    bool keep_local_block = false)
    {
        static_assert(sizeof(I) == 4, "the halo exchange indexes its send buffer with int32");
        if (!col_part) col_part = row_part;
        const int P = comm->size(), rank = comm->rank();
        if (row_part->get_num_parts() != P)
            throw DimensionMismatch("read_distributed: one part per rank of the communicator");
        auto a = assemble_local<V, I, G>(exec, data, row_part, col_part, rank, keep_local_block);
        const auto& recv_counts = a.imap->get_remote_sizes();
        std::vector<int64> send_counts(P, 0);
        array<int32> send_idx(exec, 0);
        if (P > 1) {
            // S[q][p]: everyone's receive counts
            array<int64> mine(exec, recv_counts), all(exec, (size_type)P * P);
            GKOB_CALL(b200_comm_allgather_bytes(exec->ctx(), comm->get(), mine.get_const_data(),
                                                all.get_data(), (int64)sizeof(int64) * P));
            const auto S = all.to_host();
            int64 max_remote = 0;
            for (int q = 0; q < P; ++q)
                max_remote = std::max(max_remote, std::accumulate(S.begin() + (size_t)q * P,
                                                                  S.begin() + (size_t)(q + 1) * P, int64(0)));
            // everyone's remote lists (local indices at their owners), padded to the longest
            array<int32> padded(exec, (size_type)max_remote), lists(exec, (size_type)max_remote * P);
            if (a.n_ghost)
                exec->copy(padded.get_data(), a.imap->get_remote_local_idxs().get_const_data(), a.n_ghost);
            GKOB_CALL(b200_comm_allgather_bytes(exec->ctx(), comm->get(), padded.get_const_data(),
                                                lists.get_data(), (int64)sizeof(int32) * max_remote));
            const auto lay = compute_send_layout(S, P, rank);
            send_counts = lay.send_counts;
            const int64 n_send = std::accumulate(send_counts.begin(), send_counts.end(), int64(0));
            send_idx = array<int32>(exec, (size_type)n_send);
            int64 at = 0;
            for (int q = 0; q < P; ++q) {
                if (send_counts[q])
                    exec->copy(send_idx.get_data() + at,
                               lists.get_const_data() + (size_t)q * max_remote + lay.source_offsets[q],
                               (size_type)send_counts[q]);
                at += send_counts[q];
            }
            exec->synchronize();
        }
        auto m = std::make_shared<Matrix>(exec, comm, std::move(a.local), a.n_ghost, send_counts, recv_counts,
                                          send_idx.get_const_data(), (int64)a.n_local_cols);
        const auto rg = a.imap->get_remote_global_idxs().to_host();
        m->non_local_to_global_.assign(rg.begin(), rg.end());
        m->local_only_ = a.local_only;
        return m;
    }

    size_type n_local() const { return local_->get_size().rows; }
    size_type n_local_cols() const { return n_local_cols_; }
    size_type n_ghost() const { return n_ghost_; }
    // global column index of every ghost entry of the extended vector (read_distributed only)
    const std::vector<int64>& get_non_local_to_global() const { return non_local_to_global_; }
    const matrix::Csr<V, I>* get_local_matrix() const { return local_.get(); }
    const LinOp* local_block() const override { return local_.get(); }
    // the square block of the owned columns (experimental::distributed::Matrix::get_local_matrix);
    // kept only when read_distributed was asked to (keep_local_block)
    std::shared_ptr<const matrix::Csr<V, I>> get_local_diagonal_block() const
    {
        if (!local_only_)
            throw NotSupported("distributed::Matrix: read_distributed(..., keep_local_block = true) keeps the "
                               "square local block a local solver needs");
        return local_only_;
    }
    std::shared_ptr<communicator> get_communicator() const { return comm_; }
    b200_halo* get_halo() const { return halo_; }
    // y_local = A x   (x_ext: owned part filled by the caller, ghosts by the exchange)
    void apply_extended(matrix::Dense<V>* x_ext, matrix::Dense<V>* y_local) const
    {
        if (apply_pipelined(x_ext, y_local)) return;
        // peer memory: the local SpMV gathers straight from the landing slots (no ghost copy; the
        // ghosts of this call are then in last_extended(), not in x_ext)
        V* in_place = nullptr;
        if (b200_halo_p2p_enabled(halo_) && x_ext->get_stride() == 1 &&
            cabi<V>::halo_exchange_inplace(exec_->ctx(), comm_->get(), halo_, x_ext->get_const_values(),
                                           &in_place) == B200_OK) {
            last_ext_ = in_place;
            auto view = matrix::Dense<V>::create_view(exec_, x_ext->get_size(), in_place, 1);
            local_->apply(view.get(), y_local);
            return;
        }
        last_ext_ = x_ext->get_values();
        GKOB_CALL(cabi<V>::halo_exchange(exec_->ctx(), comm_->get(), halo_, x_ext->get_values(),
                                         nullptr));
        local_->apply(x_ext, y_local);
    }
    // the extended vector [owned | ghosts] the last apply_extended gathered from
    const V* last_extended() const { return last_ext_; }
    // 1 once apply_extended runs the pipelined exchange (B200_DIST_OVERLAP=1 / set_overlap(true) and a
    // qualifying halo)
    bool pipelined() const { return owner_plan_ != nullptr && overlap_want_ == 1 && !overlap_failed_; }
    void set_overlap(bool on) const { overlap_want_ = on ? 1 : 0; }

protected:
    // Opt-in (B200_DIST_OVERLAP=1): exchange and SpMV pipelined by OWNER BLOCK.  The local matrix is kept
    // a second time split by the owner of the column ([owned | ghosts of rank 0 | ...] are contiguous
    // column ranges of the local numbering); the push runs on its own stream in ring order, and this
    // rank applies its own block first, then the block of rank-1, rank-2, ... as each lands (the SpMV of
    // a block waits for that block's arrival flag itself).  The reference overlaps the local block with
    // the exchange the same way (core/distributed/matrix.cpp:450-509: local_mtx_ / non_local_mtx_).
    // Row sums are associated in ARRIVAL order, not left to right: results agree with the single-GPU
    // ones to rounding (1e-13 relative), not bit for bit -- hence opt-in.
    bool apply_pipelined(matrix::Dense<V>* x_ext, matrix::Dense<V>* y_local) const
    {
        if (overlap_want_ < 0)
            overlap_want_ = (getenv("B200_DIST_OVERLAP") && atoi(getenv("B200_DIST_OVERLAP")) != 0) ? 1 : 0;
        const int P = comm_->size(), r = comm_->rank();
        if (overlap_want_ != 1 || overlap_failed_ || P < 2 || P > 16 || !b200_halo_p2p_enabled(halo_) ||
            x_ext->get_stride() != 1 || y_local->get_size().cols != 1)
            return false;
        auto ctx = exec_->ctx();
        std::vector<int64> recv(P, 0);
        if (!owner_plan_) {
            GKOB_CALL(b200_halo_counts(halo_, recv.data(), nullptr));
            std::vector<int64> bounds(P + 1, 0);
            bounds[1] = (int64)n_local_cols_;
            int k = 1;
            for (int q = 0; q < P; ++q) {
                if (q == r) continue;
                bounds[k + 1] = bounds[k] + recv[q];
                ++k;
            }
            const auto nnz = local_->get_num_stored_elements();
            b200_csr_plan* p = nullptr;
            GKOB_CALL((viabi<V, I>::csr_plan_create(ctx, local_->get_size().rows, nnz, local_->get_const_row_ptrs(), &p)));
            b200_csr_plan_allow_value_copy(p, 1);
            b200_status st = viabi<V, I>::csr_plan_tune(ctx, p, local_->get_size().rows, local_->get_size().cols, nnz,
                                                       local_->get_const_row_ptrs(), local_->get_const_col_idxs(),
                                                       local_->get_const_values());
            if (st == B200_OK)
                st = viabi<V, I>::csr_plan_split_columns(ctx, p, local_->get_size().rows, local_->get_size().cols, nnz,
                                                         local_->get_const_row_ptrs(), local_->get_const_col_idxs(),
                                                         local_->get_const_values(), P, bounds.data());
            if (st != B200_OK || b200_csr_plan_parts(p) != P) {
                std::fprintf(stderr, "[gko_b200] rank %d: pipelined exchange unavailable: owner split gave %d parts, "
                                     "status %d (%s)\n",
                             r, (int)b200_csr_plan_parts(p), (int)st, b200_last_error());
                b200_csr_plan_destroy(p);
                overlap_failed_ = true;
                return false;
            }
            owner_plan_ = p;
            recv_counts_ = recv;
        }
        V* b = nullptr;
        const uint64_t* flags = nullptr;
        uint64_t epoch = 0;
        if (cabi<V>::halo_exchange_staged_begin(ctx, comm_->get(), halo_, x_ext->get_const_values(), &b, &flags,
                                                &epoch) != B200_OK) {
            std::fprintf(stderr, "[gko_b200] rank %d: pipelined exchange unavailable: %s\n", r, b200_last_error());
            overlap_failed_ = true;
            return false;
        }
        for (int s = 0; s < P; ++s) {
            const int src = (r - s + P) % P;
            if (s > 0 && recv_counts_[src] == 0) continue;  // nothing comes from there: no flag either
            const int part = s == 0 ? 0 : (src < r ? src + 1 : src);
            GKOB_CALL((viabi<V, I>::csr_spmv_part(ctx, owner_plan_, part, s > 0 ? 1 : 0, b, 1, y_local->get_values(),
                                                  y_local->get_stride(), s == 0 ? nullptr : flags + src, epoch)));
        }
        GKOB_CALL(b200_halo_exchange_staged_end(ctx, halo_));
        last_ext_ = b;
        return true;
    }

public:
    // a distributed::Vector of this matrix's row / column partition
    std::unique_ptr<Vector<V>> create_row_vector(size_type global_rows) const
    {
        return Vector<V>::create(exec_, comm_, dim2{global_rows, 1}, dim2{n_local(), 1});
    }

protected:
    // LinOp::apply on the LOCAL rows of distributed vectors: column by column, b's rows are
    // copied next to the ghost slots of an internal extended vector, then exchange + local SpMV
    // (experimental::distributed::Matrix::apply_impl, core/distributed/matrix.cpp:450-509)
    matrix::Dense<V>* gather(const matrix::Dense<V>* b, size_type col) const
    {
        if (!x_ext_) x_ext_ = matrix::Dense<V>::create(exec_, dim2{n_local_cols_ + n_ghost_, 1});
        auto owned = x_ext_->create_submatrix_rows(0, n_local_cols_);
        auto bcol = matrix::Dense<V>::create_view(exec_, dim2{b->get_size().rows, 1},
                                                  const_cast<V*>(b->get_const_values()) + col, b->get_stride());
        owned->copy_from(bcol.get());
        GKOB_CALL(cabi<V>::halo_exchange(exec_->ctx(), comm_->get(), halo_, x_ext_->get_values(), nullptr));
        return x_ext_.get();
    }
    static std::unique_ptr<matrix::Dense<V>> column_of(matrix::Dense<V>* x, size_type col)
    {
        return matrix::Dense<V>::create_view(x->get_executor(), dim2{x->get_size().rows, 1},
                                             x->get_values() + col, x->get_stride());
    }
    void apply_impl(const LinOp* lb, LinOp* lx) const override
    {
        auto b = as<matrix::Dense<V>>(lb);
        auto x = as<matrix::Dense<V>>(lx);
        for (size_type j = 0; j < b->get_size().cols; ++j) {
            auto xe = gather(b, j);
            local_->apply(xe, column_of(x, j).get());
        }
    }
    void apply_impl(const LinOp* alpha, const LinOp* lb, const LinOp* beta, LinOp* lx) const override
    {
        auto b = as<matrix::Dense<V>>(lb);
        auto x = as<matrix::Dense<V>>(lx);
        for (size_type j = 0; j < b->get_size().cols; ++j) {
            auto xe = gather(b, j);
            local_->apply(alpha, xe, beta, column_of(x, j).get());
        }
    }

private:
    mutable std::unique_ptr<matrix::Dense<V>> x_ext_;
    mutable const V* last_ext_ = nullptr;
    mutable b200_csr_plan* owner_plan_ = nullptr;  // the local matrix split by column owner (pipelined apply)
    mutable std::vector<int64> recv_counts_;
    mutable bool overlap_failed_ = false;
    mutable int overlap_want_ = -1;  // -1: take B200_DIST_OVERLAP
    std::shared_ptr<communicator> comm_;
    std::unique_ptr<matrix::Csr<V, I>> local_;
    size_type n_ghost_;
    size_type n_local_cols_ = 0;
    b200_halo* halo_ = nullptr;
    std::vector<int64> non_local_to_global_;  // read_distributed: global index of every ghost column
    std::shared_ptr<matrix::Csr<V, I>> local_only_;
};

namespace preconditioner {

// experimental::distributed::preconditioner::Schwarz, one level (core/distributed/preconditioner/
// schwarz.cpp:88-140 without the coarse correction and the L1 smoother): a local solver,
// generated from the square local block of the distributed matrix, applied to the local rows.
template <typename V, typename I>
class Schwarz : public LinOp {
public:
    struct Factory : LinOpFactory {
        std::shared_ptr<const LinOpFactory> local_solver_;
        std::shared_ptr<const LinOp> generated_local_solver_;
        std::shared_ptr<const Executor> exec_;
        Factory& with_local_solver(std::shared_ptr<const LinOpFactory> f)
        {
            local_solver_ = std::move(f);
            return *this;
        }
        Factory& with_generated_local_solver(std::shared_ptr<const LinOp> s)
        {
            generated_local_solver_ = std::move(s);
            return *this;
        }
        std::shared_ptr<const LinOpFactory> on(std::shared_ptr<const Executor> exec) const
        {
            auto f = std::make_shared<Factory>(*this);
            f->exec_ = exec;
            return f;
        }
        std::unique_ptr<LinOp> generate(std::shared_ptr<const LinOp> op) const override
        {
            return std::unique_ptr<LinOp>(new Schwarz(exec_ ? exec_ : op->get_executor(), *this, op));
        }
    };
    static Factory build() { return Factory{}; }
    std::shared_ptr<const LinOp> get_local_solver() const { return local_solver_; }
    bool apply_uses_initial_guess() const override { return local_solver_->apply_uses_initial_guess(); }

protected:
    Schwarz(std::shared_ptr<const Executor> exec, const Factory& f, std::shared_ptr<const LinOp> op)
        : LinOp(exec, dim2{op->get_size().rows, op->get_size().rows})
    {
        if (f.generated_local_solver_) {
            local_solver_ = f.generated_local_solver_;
        } else {
            if (!f.local_solver_) throw NotSupported("Schwarz: a local solver (factory) is required");
            auto dist = dynamic_cast<const Matrix<V, I>*>(op.get());
            if (!dist) throw NotSupported("Schwarz: the operator is not a distributed::Matrix");
            local_solver_ = f.local_solver_->generate(dist->get_local_diagonal_block());
        }
        if (local_solver_->get_size().rows != size_.rows)
            throw DimensionMismatch("Schwarz: local solver and local rows differ in size");
    }
    // the local rows WITHOUT the sum over the ranks (gko::detail::get_local): a local solver's dots
    // and norms are local
    static std::unique_ptr<matrix::Dense<V>> local_view(const LinOp* v)
    {
        auto d = as<matrix::Dense<V>>(v);
        return matrix::Dense<V>::create_view(d->get_executor(), d->get_size(),
                                             const_cast<V*>(d->get_const_values()), d->get_stride());
    }
    void apply_impl(const LinOp* b, LinOp* x) const override
    {
        auto lb = local_view(b), lx = local_view(x);
        local_solver_->apply(lb.get(), lx.get());
    }
    void apply_impl(const LinOp* alpha, const LinOp* b, const LinOp* beta, LinOp* x) const override
    {
        auto lb = local_view(b), lx = local_view(x);
        local_solver_->apply(alpha, lb.get(), beta, lx.get());
    }

private:
    std::shared_ptr<const LinOp> local_solver_;
};

}  // namespace preconditioner

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
        // After the stop (ctl[0] != 0) the remaining iterations of a check_every_ batch are no-op
        // KERNELS, but their all-reduces still run on the already-global sc[2] and sc[6..7]: those
        // three cells are scratch between two finishes and UNDEFINED after a stop (they are re-summed
        // once per left-over iteration).  Nothing reads them then: the results of the solve are
        // x, ctl[0..1] and sc[0], sc[1], sc[3..5], which the all-reduces never touch.
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
        if (h[0] == 0) {
            // One batch of iterations is always queued AHEAD of the one the host looks at: the control
            // block is read through a stream-ordered snapshot taken at the batch boundary
            // (b200_snapshot_*), so neither the GPU nor -- distributed -- the other ranks' GPUs wait for
            // this host's round trip, and every rank sees the state of the same boundary.  After the
            // stop the kernels of the batch queued ahead are no-ops.
            int slot = 0;
            int32 before = h[1];
            GKOB_CALL(b200_graph_launch(ctx, graph_));
            GKOB_CALL(b200_snapshot_begin(ctx, slot, ctl_.get_const_data(), sizeof(h)));
            for (;;) {
                GKOB_CALL(b200_graph_launch(ctx, graph_));
                GKOB_CALL(b200_snapshot_begin(ctx, slot ^ 1, ctl_.get_const_data(), sizeof(h)));
                GKOB_CALL(b200_snapshot_end(ctx, slot, h, sizeof(h)));
                if (h[0] != 0) break;
                if (h[1] == before) throw Error("fused CG: the device iteration made no progress");
                before = h[1];
                slot ^= 1;
            }
            exec_->copy_to_host(h, ctl_.get_const_data(), 8);  // drains the batch queued ahead
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
