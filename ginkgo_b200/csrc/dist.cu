// Multi-GPU support: 1-D row partition over the GPUs of one box, NCCL over NVLink 5 /
// NVSwitch.  Mirrors the design of the reference's experimental::distributed::Matrix
// (core/distributed/matrix.cpp:450-509: local block + non-local block with COMPRESSED
// column numbering + a row gatherer that exchanges only the referenced remote entries)
// and distributed::Vector reductions (core/distributed/vector.cpp:510-534: local
// reduction + all-reduce), with MPI replaced by NCCL on the compute stream (graph
// capturable, no host staging, no exec->synchronize() before the collective).
//
// Layout on rank p: rows [offsets[p], offsets[p+1]).  Vectors that feed an SpMV are
// "extended": [ n_local owned entries | n_ghost received entries ], ghosts ordered by
// global column (hence grouped by owning rank).  The local CSR keeps every row's entries
// in their original order, only the column numbers are remapped into the extended
// vector -- so the row sums are bit-identical to the single-GPU result.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded in
// the process, or the system one), so the library itself has no link-time dependency.
#include <dlfcn.h>

#include <vector>

#include "common.cuh"

namespace b200 {
namespace nccl {

// minimal NCCL ABI (nccl.h is stable across 2.x for these entry points)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt64 = 4, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Api {
    int (*GetUniqueId)(ncclUniqueId*);
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
    bool ok = false;
};

static Api& api()
{
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return a;
    }
#define B200_SYM(name)                                       \
    *(void**)(&a.name) = dlsym(h, "nccl" #name);             \
    if (!a.name) {                                           \
        set_error("libnccl lacks nccl" #name);               \
        return a;                                            \
    }
    B200_SYM(GetUniqueId)
    B200_SYM(CommInitRank)
    B200_SYM(CommDestroy)
    B200_SYM(AllReduce)
    B200_SYM(AllGather)
    B200_SYM(Send)
    B200_SYM(Recv)
    B200_SYM(GroupStart)
    B200_SYM(GroupEnd)
    B200_SYM(GetErrorString)
#undef B200_SYM
    a.ok = true;
    return a;
}

template <typename T>
struct dtype;
template <>
struct dtype<double> {
    static constexpr int v = ncclFloat64;
};
template <>
struct dtype<float> {
    static constexpr int v = ncclFloat32;
};
template <>
struct dtype<int32_t> {
    static constexpr int v = ncclInt32;
};
template <>
struct dtype<int64_t> {
    static constexpr int v = ncclInt64;
};

}  // namespace nccl
}  // namespace b200

#define B200_NCCL_CHECK(expr)                                                               \
    do {                                                                                    \
        int r__ = (expr);                                                                   \
        if (r__ != b200::nccl::ncclSuccess) {                                               \
            b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                            b200::nccl::api().GetErrorString(r__));                         \
            return B200_ERR_COMM;                                                           \
        }                                                                                   \
    } while (0)

struct b200_comm {
    b200::nccl::ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

// halo plan: which owned entries go to which peer, where received entries land
struct b200_halo {
    int nranks = 1;
    int64_t n_local = 0, n_ghost = 0, n_send = 0;
    std::vector<int64_t> send_count, send_off, recv_count, recv_off;  // per peer
    int32_t* send_idx = nullptr;  // device: local indices to pack, grouped by peer
    void* send_buf = nullptr;     // device: n_send values
    size_t elem = 0;
    int device = 0;
};

namespace b200 {
namespace dist {

template <typename V>
__global__ void pack_kernel(int64_t n, const int32_t* __restrict__ idx, const V* __restrict__ x,
                            V* __restrict__ out, const int32_t* __restrict__ ctl)
{
    if (ctl && ctl[0] != 0) return;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[idx[i]];
}

}  // namespace dist
}  // namespace b200

extern "C" {

b200_status b200_comm_get_unique_id(uint8_t* id128)
{
    auto& a = b200::nccl::api();
    if (!a.ok) return B200_ERR_COMM;
    b200::nccl::ncclUniqueId id;
    B200_NCCL_CHECK(a.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return B200_OK;
}

b200_status b200_comm_create(b200_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t nranks,
                             b200_comm** out)
{
    B200_REQUIRE(ctx && id128 && out, "null argument");
    auto& a = b200::nccl::api();
    if (!a.ok) return B200_ERR_COMM;
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    b200::nccl::ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    b200_comm* c = new b200_comm();
    c->rank = rank;
    c->nranks = nranks;
    int r = a.CommInitRank(&c->comm, nranks, id, rank);
    if (r != b200::nccl::ncclSuccess) {
        b200::set_error("ncclCommInitRank: %s", a.GetErrorString(r));
        delete c;
        return B200_ERR_COMM;
    }
    *out = c;
    return B200_OK;
}

void b200_comm_destroy(b200_comm* comm)
{
    if (!comm) return;
    if (comm->comm) b200::nccl::api().CommDestroy(comm->comm);
    delete comm;
}

int32_t b200_comm_rank(const b200_comm* c) { return c->rank; }
int32_t b200_comm_size(const b200_comm* c) { return c->nranks; }

void b200_halo_destroy(b200_halo* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->send_idx);
    cudaFree(h->send_buf);
    delete h;
}
int64_t b200_halo_num_ghost(const b200_halo* h) { return h->n_ghost; }
int64_t b200_halo_num_send(const b200_halo* h) { return h->n_send; }

#define B200_DEF_COMM(V, VT)                                                                   \
    /* in-place sum all-reduce of `count` device values, enqueued on the ctx stream */         \
    b200_status b200_comm_allreduce_sum_##V(b200_ctx* ctx, b200_comm* comm, VT* buf,           \
                                            int64_t count)                                     \
    {                                                                                          \
        auto& a = b200::nccl::api();                                                           \
        if (comm->nranks == 1) return B200_OK;                                                 \
        B200_NCCL_CHECK(a.AllReduce(buf, buf, (size_t)count, b200::nccl::dtype<VT>::v,         \
                                    b200::nccl::ncclSum, comm->comm, ctx->stream));            \
        return B200_OK;                                                                        \
    }                                                                                          \
    b200_status b200_comm_allgather_##V(b200_ctx* ctx, b200_comm* comm, const VT* send,        \
                                        VT* recv, int64_t count_per_rank)                      \
    {                                                                                          \
        auto& a = b200::nccl::api();                                                           \
        B200_NCCL_CHECK(a.AllGather(send, recv, (size_t)count_per_rank,                        \
                                    b200::nccl::dtype<VT>::v, comm->comm, ctx->stream));       \
        return B200_OK;                                                                        \
    }                                                                                          \
    /* x_ext = [n_local owned | n_ghost]: pack the entries peers need, exchange, ghosts land */\
    /* in the tail of x_ext.  ctl (optional): fused-solver stop flag (the NCCL calls still  */ \
    /* run -- every rank takes the same decision, they carry stale data harmlessly).        */ \
    b200_status b200_halo_exchange_##V(b200_ctx* ctx, b200_comm* comm, b200_halo* h,           \
                                       VT* x_ext, const int32_t* ctl)                          \
    {                                                                                          \
        auto& a = b200::nccl::api();                                                           \
        if (h->nranks == 1 || (h->n_send == 0 && h->n_ghost == 0)) return B200_OK;             \
        VT* sb = (VT*)h->send_buf;                                                             \
        if (h->n_send > 0) {                                                                   \
            b200::dist::pack_kernel<VT>                                                        \
                <<<(unsigned)b200::ceildiv(h->n_send, 256), 256, 0, ctx->stream>>>(            \
                    h->n_send, h->send_idx, x_ext, sb, ctl);                                   \
            B200_LAUNCH_CHECK(ctx);                                                            \
        }                                                                                      \
        B200_NCCL_CHECK(a.GroupStart());                                                       \
        for (int p = 0; p < h->nranks; ++p) {                                                  \
            if (h->send_count[p] > 0)                                                          \
                B200_NCCL_CHECK(a.Send(sb + h->send_off[p], (size_t)h->send_count[p],          \
                                       b200::nccl::dtype<VT>::v, p, comm->comm, ctx->stream)); \
            if (h->recv_count[p] > 0)                                                          \
                B200_NCCL_CHECK(a.Recv(x_ext + h->n_local + h->recv_off[p],                    \
                                       (size_t)h->recv_count[p], b200::nccl::dtype<VT>::v, p,  \
                                       comm->comm, ctx->stream));                              \
        }                                                                                      \
        B200_NCCL_CHECK(a.GroupEnd());                                                         \
        return B200_OK;                                                                        \
    }
B200_DEF_COMM(f64, double)
B200_DEF_COMM(f32, float)

// Halo plan from the partition set-up (done by the host side with torch.distributed, see
// ginkgo_b200/distributed.py): per-peer counts (host arrays of nranks entries) and the
// device list of owned entries to pack, grouped by destination rank.
b200_status b200_halo_create(b200_ctx* ctx, int32_t nranks, int64_t n_local, int64_t n_ghost,
                             const int64_t* send_counts, const int64_t* recv_counts,
                             const int32_t* send_idx_dev, int32_t value_bytes, b200_halo** out)
{
    B200_REQUIRE(ctx && send_counts && recv_counts && out, "null argument");
    b200_halo* h = new b200_halo();
    h->nranks = nranks;
    h->n_local = n_local;
    h->n_ghost = n_ghost;
    h->elem = value_bytes;
    h->device = ctx->device;
    h->send_count.assign(send_counts, send_counts + nranks);
    h->recv_count.assign(recv_counts, recv_counts + nranks);
    h->send_off.assign(nranks, 0);
    h->recv_off.assign(nranks, 0);
    int64_t ns = 0, nr = 0;
    for (int p = 0; p < nranks; ++p) {
        h->send_off[p] = ns;
        h->recv_off[p] = nr;
        ns += h->send_count[p];
        nr += h->recv_count[p];
    }
    h->n_send = ns;
    if (nr != n_ghost) {
        delete h;
        b200::set_error("halo: recv counts do not add up to n_ghost");
        return B200_ERR_INVALID;
    }
    if (ns > 0) {
        B200_CUDA_CHECK(cudaMalloc((void**)&h->send_idx, sizeof(int32_t) * ns));
        B200_CUDA_CHECK(cudaMalloc(&h->send_buf, (size_t)value_bytes * ns));
        B200_CUDA_CHECK(cudaMemcpyAsync(h->send_idx, send_idx_dev, sizeof(int32_t) * ns,
                                        cudaMemcpyDeviceToDevice, ctx->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    *out = h;
    return B200_OK;
}

}  // extern "C"
