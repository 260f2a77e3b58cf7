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

// A peer window: one cudaMalloc'd block per rank, exported with cudaIpcGetMemHandle and
// mapped by every other rank of the box, so kernels can store straight into a peer's HBM
// through NVLink / NVSwitch.
struct b200_peer_window {
    void* local = nullptr;
    size_t bytes = 0;
    std::vector<void*> mapped;  // per rank, mapped[rank] == local
};

struct b200_comm {
    b200::nccl::ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory all-reduce of a few scalars (b200_comm_enable_p2p)
    bool p2p = false;
    b200_peer_window win;           // [flags 2 x nranks u64 | mailbox 2 x nranks x 8 cells of 8 B]
    void** peer_base_dev = nullptr;  // device: nranks window base pointers
    uint64_t* epoch_dev = nullptr;   // device: number of completed all-reduces
    int* err_dev = nullptr;          // device: set by a kernel whose wait timed out (sticky)
    int device = 0;
    b200_ctx* ctx = nullptr;         // the context enable_p2p ran on (must outlive the comm)
    // exported blocks of destroyed halos: freed in b200_comm_destroy after a barrier, because
    // an exported block must not be freed while a peer still has it mapped
    std::vector<void*> retired;
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
    // peer-memory exchange (b200_halo_enable_p2p): entries are stored straight into the
    // owner-side landing slots of the peers, no NCCL call on the data path
    bool p2p = false;
    int rank = 0;
    // window: [flags 2 x nranks u64 | extended vector 0 | extended vector 1]; an extended vector is
    // [n_local owned | n_ghost landing slot], so a kernel can gather straight from it
    // (b200_halo_exchange_inplace_*) -- or the slot is copied into the caller's ghost tail
    b200_peer_window win;
    size_t buf_off[2] = {0, 0};     // the two extended vectors
    size_t slot_off[2] = {0, 0};    // their ghost tails (= landing slots)
    // every peer's send list is one contiguous run of owned entries (a matrix whose ghosts are
    // (nearly) all of x, e.g. uniformly random columns): pushed with 16-byte remote stores
    bool runs = false;
    int64_t* run_base_dev = nullptr;  // device: first owned index of the run, per peer
    // pipelined exchange (b200_halo_exchange_staged_*): the push runs on its own stream, destinations
    // in ring order (rank + 1, rank + 2, ...), one arrival flag per source
    cudaStream_t push_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_pushed = nullptr;
    unsigned* dst_ticket_dev = nullptr;  // device [nranks], zero, self-resetting
    uint64_t host_epoch = 0;          // exchanges issued from the host outside stream capture
    bool captured = false;            // an exchange was captured into a graph: host_epoch is unreliable
    int64_t* meta_dev = nullptr;    // device: send_off[nranks+1] | send_cnt[nranks] | recv_cnt[nranks]
    void** peer_slot_dev = nullptr;  // device: [2][nranks] where my segment starts on peer p
    uint64_t** peer_flag_dev = nullptr;  // device: [nranks] &flags[0][rank] on peer p
    uint64_t* epoch_dev = nullptr;
    unsigned* ticket_dev = nullptr;  // 2 self-resetting counters
    int* err_dev = nullptr;          // the communicator's error word
    b200_comm* comm = nullptr;       // owner of the retired-block list (must outlive the halo)
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


// ---------------------------------------------------------------------------------------
// Peer-memory collectives.  Synchronisation is by epoch-stamped flags in the receiver's
// window: a sender stores its data, fences at system scope and releases flag[parity][sender]
// = epoch on the receiver; the receiver acquires the flag before it reads.  Two parities:
// a rank can only start epoch e+2 after every peer it talks to has finished reading epoch e
// (it had to receive their epoch e+1 first), so slot e & 1 is free again.  The epoch lives
// in device memory and is advanced by the kernels themselves, which keeps every call
// capturable in a CUDA graph.  A wait that exceeds kSpinLimit cycles sets a sticky error
// word instead of hanging the GPU.
// ---------------------------------------------------------------------------------------
constexpr long long kSpinLimit = 20000000000ll;  // ~10 s of SM clock

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag(const uint64_t* p, uint64_t epoch, int* err)
{
    if (*(volatile int*)err) return;
    const long long t0 = clock64();
    int polls = 0;
    while (ld_acquire_sys(p) < epoch) {
        if (++polls > 4096) {  // back off only once the wait is clearly not a short one
            if (clock64() - t0 > kSpinLimit) {
                *(volatile int*)err = 1;
                return;
            }
            __nanosleep(100);
        }
    }
}

struct HaloDev {
    int nranks, rank;
    int64_t n_local, n_ghost, n_send;
    const int32_t* send_idx;
    const int64_t* meta;       // send_off[nranks+1] | send_cnt[nranks] | recv_cnt[nranks]
    void* const* peer_slot;    // [2][nranks]
    uint64_t* const* peer_flag;  // [nranks]
    uint64_t* my_flags;        // [2][nranks]
    const void* my_slot[2];
    uint64_t* epoch;
    unsigned* ticket;
    int* err;
};

// pack + send in one kernel: every owned entry a peer references is stored directly into
// that peer's landing slot (remote store over NVLink); the last CTA to finish releases the
// epoch flag on every peer that receives from this rank.
template <typename V>
__global__ void __launch_bounds__(256) halo_push_kernel(HaloDev h, const V* __restrict__ x)
{
    const uint64_t e = *(volatile uint64_t*)h.epoch + 1;
    const int par = (int)(e & 1);
    const int64_t* send_off = h.meta;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = 4;  // independent gather -> remote-store chains per thread
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < h.n_send;
         i0 += U * stride) {
        int32_t idx[U];
        V v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int64_t i = i0 + k * stride;
            idx[k] = i < h.n_send ? h.send_idx[i] : 0;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = x[idx[k]];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int64_t i = i0 + k * stride;
            if (i < h.n_send) {
                int p = 0;
                while (i >= send_off[p + 1]) ++p;
                V* dst = (V*)h.peer_slot[par * h.nranks + p];
                dst[i - send_off[p]] = v[k];
            }
        }
    }
    // one system-scope fence per CTA: the barrier orders every thread's stores before
    // thread 0's fence, which is cumulative (the cooperative-groups grid-sync idiom)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(h.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *h.ticket = 0;
        const int64_t* send_cnt = h.meta + h.nranks + 1;
        if ((int)threadIdx.x < h.nranks && send_cnt[threadIdx.x] > 0) {
            __threadfence_system();
            st_release_sys(h.peer_flag[threadIdx.x] + par * h.nranks, e);
        }
    }
}

// wait + unpack: acquire the epoch flag of every rank this one receives from, then move the
// landing slot into the ghost tail of the extended vector.  The last CTA advances the epoch.
template <typename V>
__global__ void __launch_bounds__(256) halo_wait_kernel(HaloDev h, V* __restrict__ x_ext)
{
    const uint64_t e = *(volatile uint64_t*)h.epoch + 1;
    const int par = (int)(e & 1);
    const int64_t* recv_cnt = h.meta + 2 * h.nranks + 1;
    if ((int)threadIdx.x < h.nranks && recv_cnt[threadIdx.x] > 0)
        wait_flag(h.my_flags + par * h.nranks + threadIdx.x, e, h.err);
    __syncthreads();
    const V* src = (const V*)h.my_slot[par];
    V* dst = x_ext + h.n_local;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < h.n_ghost; i += stride)
        dst[i] = __ldcg(src + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(h.ticket + 1, 1u) == gridDim.x - 1) {
            h.ticket[1] = 0;
            *(volatile uint64_t*)h.epoch = e;
        }
    }
}

// Contiguous runs: peer p receives x[run_base[p] .. + send_cnt[p]).  The work of all destinations
// (plus, with own != nullptr, the copy of the owned entries into this rank's own extended vector,
// so the consumer kernel can read [owned | ghosts] from one base pointer) is cut into blocks of
// kRunBlock elements dealt round-robin to the CTAs; a thread keeps 4 independent 16-byte stores in
// flight (two 8-byte local loads each), 8-byte stores only at unaligned segment ends.  First
// version: 32 CTAs per destination, one load -> store chain per thread = 278 GB/s on one NVLink
// peer (profiles/r02i_multi_gpu.txt); this one keeps every link busy.
constexpr int kRunBlock = 2048;

template <typename V>
__global__ void __launch_bounds__(256) halo_push_runs_kernel(HaloDev h, const V* __restrict__ x,
                                                            const int64_t* __restrict__ run_base, V* own)
{
    const uint64_t e = *(volatile uint64_t*)h.epoch + 1;
    const int par = (int)(e & 1);
    const int64_t* send_cnt = h.meta + h.nranks + 1;
    const int ndst = h.nranks + (own ? 1 : 0);
    // blocks per destination (destination nranks = the own copy)
    int64_t total = 0;
    for (int p = 0; p < ndst; ++p) total += ceildiv(p < h.nranks ? send_cnt[p] : h.n_local, (int64_t)kRunBlock);
    constexpr int kPer = 16 / (int)sizeof(V);  // elements per 16-byte store
    for (int64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
        int p = 0;
        int64_t first = 0;
        for (;; ++p) {
            const int64_t nb = ceildiv(p < h.nranks ? send_cnt[p] : h.n_local, (int64_t)kRunBlock);
            if (blk < first + nb) break;
            first += nb;
        }
        const int64_t cnt = p < h.nranks ? send_cnt[p] : h.n_local;
        const V* src = p < h.nranks ? x + run_base[p] : x;
        V* dst = p < h.nranks ? (V*)h.peer_slot[par * h.nranks + p] : own;
        const int64_t lo = (blk - first) * kRunBlock;
        int64_t hi = lo + kRunBlock;
        if (hi > cnt) hi = cnt;
        // 16-byte aligned body of [lo, hi) in destination space
        int64_t head = (int64_t)(((16 - ((uintptr_t)(dst + lo) & 15)) & 15) / sizeof(V));
        if (head > hi - lo) head = hi - lo;
        const int64_t body = (hi - lo - head) / kPer;
        const int tid = threadIdx.x;
        if (tid < head) dst[lo + tid] = src[lo + tid];
        const V* s2 = src + lo + head;
        V* d2 = dst + lo + head;
        // kRunBlock / kPer <= 1024 stores per block = 4 per thread, all loads first
        V v[4][kPer];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = tid + 256 * u;
            if (i < body) {
#pragma unroll
                for (int k = 0; k < kPer; ++k) v[u][k] = s2[i * kPer + k];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = tid + 256 * u;
            if (i < body) {
                if (sizeof(V) == 8) {
                    *reinterpret_cast<double2*>(d2 + i * kPer) = make_double2((double)v[u][0], (double)v[u][kPer - 1]);
                } else {
                    *reinterpret_cast<float4*>(d2 + i * kPer) =
                        make_float4((float)v[u][0], (float)v[u][1 % kPer], (float)v[u][2 % kPer], (float)v[u][3 % kPer]);
                }
            }
        }
        const int64_t done = lo + head + body * kPer;
        if (tid < hi - done) dst[done + tid] = src[done + tid];
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(h.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *h.ticket = 0;
        if ((int)threadIdx.x < h.nranks && send_cnt[threadIdx.x] > 0) {
            __threadfence_system();
            st_release_sys(h.peer_flag[threadIdx.x] + par * h.nranks, e);
        }
    }
}

// Pipelined variant of the run push: destinations in RING order (rank + 1, rank + 2, ...; every rank
// sends to a different peer in every step, so the first block a rank needs -- from rank - 1 -- is
// complete after 1 / (P - 1) of the exchange), blocks dealt step-major, and the CTA that completes a
// destination releases that destination's flag at once.  Runs on its own stream next to the SpMV
// launches that consume the blocks in arrival order (few CTAs: NVLink is saturated by ~100).
template <typename V>
__global__ void __launch_bounds__(256) halo_push_staged_kernel(HaloDev h, const V* __restrict__ x,
                                                              const int64_t* __restrict__ run_base,
                                                              unsigned* __restrict__ dst_ticket)
{
    const uint64_t e = *(volatile uint64_t*)h.epoch + 1;
    const int par = (int)(e & 1);
    const int64_t* send_cnt = h.meta + h.nranks + 1;
    int64_t total = 0;
    for (int p = 0; p < h.nranks; ++p) total += ceildiv(send_cnt[p], (int64_t)kRunBlock);
    constexpr int kPer = 16 / (int)sizeof(V);
    __shared__ bool last;
    for (int64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
        int p = 0;
        int64_t first = 0, nb = 0;
        for (int s = 1; s < h.nranks; ++s) {  // ring order
            p = (h.rank + s) % h.nranks;
            nb = ceildiv(send_cnt[p], (int64_t)kRunBlock);
            if (blk < first + nb) break;
            first += nb;
        }
        const int64_t cnt = send_cnt[p];
        const V* src = x + run_base[p];
        V* dst = (V*)h.peer_slot[par * h.nranks + p];
        const int64_t lo = (blk - first) * kRunBlock;
        int64_t hi = lo + kRunBlock;
        if (hi > cnt) hi = cnt;
        int64_t head = (int64_t)(((16 - ((uintptr_t)(dst + lo) & 15)) & 15) / sizeof(V));
        if (head > hi - lo) head = hi - lo;
        const int64_t body = (hi - lo - head) / kPer;
        const int tid = threadIdx.x;
        if (tid < head) dst[lo + tid] = src[lo + tid];
        const V* s2 = src + lo + head;
        V* d2 = dst + lo + head;
        V v[4][kPer];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = tid + 256 * u;
            if (i < body) {
#pragma unroll
                for (int k = 0; k < kPer; ++k) v[u][k] = s2[i * kPer + k];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = tid + 256 * u;
            if (i < body) {
                if (sizeof(V) == 8) {
                    *reinterpret_cast<double2*>(d2 + i * kPer) = make_double2((double)v[u][0], (double)v[u][kPer - 1]);
                } else {
                    *reinterpret_cast<float4*>(d2 + i * kPer) =
                        make_float4((float)v[u][0], (float)v[u][1 % kPer], (float)v[u][2 % kPer], (float)v[u][3 % kPer]);
                }
            }
        }
        const int64_t done = lo + head + body * kPer;
        if (tid < hi - done) dst[done + tid] = src[done + tid];
        // destination p complete?  (the CTA that stores its last block releases p's flag)
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            last = atomicAdd(dst_ticket + p, 1u) == (unsigned)nb - 1;
            if (last) {
                dst_ticket[p] = 0;
                __threadfence_system();
                st_release_sys(h.peer_flag[p] + par * h.nranks, e);
            }
        }
        __syncthreads();
    }
}

__global__ void halo_epoch_advance_kernel(HaloDev h)
{
    *(volatile uint64_t*)h.epoch = *(volatile uint64_t*)h.epoch + 1;
}

// wait only: acquire the epoch flag of every rank this one receives from; the consumer reads the
// landing slot in place.  One CTA; advances the epoch.
__global__ void __launch_bounds__(256) halo_wait_inplace_kernel(HaloDev h)
{
    const uint64_t e = *(volatile uint64_t*)h.epoch + 1;
    const int par = (int)(e & 1);
    const int64_t* recv_cnt = h.meta + 2 * h.nranks + 1;
    if ((int)threadIdx.x < h.nranks && recv_cnt[threadIdx.x] > 0)
        wait_flag(h.my_flags + par * h.nranks + threadIdx.x, e, h.err);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        *(volatile uint64_t*)h.epoch = e;
    }
}

struct CommDev {
    int nranks, rank;
    void* const* peer_base;  // [nranks] window bases
    uint64_t* epoch;
    int* err;
};
constexpr int kMailCells = 8;  // values per all-reduce, 8-byte cells

// sum all-reduce of <= 8 scalars in ONE single-CTA kernel: thread p stores this rank's values
// into rank p's mailbox and releases its flag there, thread q waits for rank q's flag here,
// then the values are added in rank order -- every rank computes the same bits, and the
// result does not depend on a ring / tree schedule.
template <typename V>
__global__ void __launch_bounds__(64) p2p_allreduce_kernel(CommDev c, V* __restrict__ buf, int count)
{
    const uint64_t e = *(volatile uint64_t*)c.epoch + 1;
    const int par = (int)(e & 1);
    const int t = threadIdx.x;
    const size_t flag_bytes = 2 * (size_t)c.nranks * sizeof(uint64_t);
    __shared__ V mine[kMailCells];
    if (t < count) mine[t] = buf[t];
    __syncthreads();
    if (t < c.nranks) {
        char* base = (char*)c.peer_base[t];
        V* cell = (V*)(base + flag_bytes + ((size_t)(par * c.nranks + c.rank) * kMailCells) * 8);
        for (int k = 0; k < count; ++k) cell[k] = mine[k];
        __threadfence_system();
        st_release_sys((uint64_t*)base + par * c.nranks + c.rank, e);
    }
    char* my = (char*)c.peer_base[c.rank];
    if (t < c.nranks) wait_flag((const uint64_t*)my + par * c.nranks + t, e, c.err);
    __syncthreads();
    if (t < count) {
        V s = V(0);
        for (int q = 0; q < c.nranks; ++q) {
            const V* cell =
                (const V*)(my + flag_bytes + ((size_t)(par * c.nranks + q) * kMailCells) * 8);
            const V v = __ldcg(cell + t);
            s = q == 0 ? v : s + v;
        }
        buf[t] = s;
    }
    __syncthreads();
    if (t == 0) *(volatile uint64_t*)c.epoch = e;
}

}  // namespace dist
}  // namespace b200

namespace {

using b200::nccl::api;

inline size_t round256(size_t b) { return (b + 255) & ~size_t(255); }

// all-gather of `bytes` host bytes per rank through the NCCL communicator (set-up only)
b200_status allgather_host(b200_ctx* ctx, b200_comm* comm, const void* mine, size_t bytes,
                           void* all)
{
    auto& a = api();
    uint8_t* dev = nullptr;
    B200_CUDA_CHECK(cudaMalloc((void**)&dev, bytes * (comm->nranks + 1)));
    B200_CUDA_CHECK(cudaMemcpyAsync(dev, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int r = a.AllGather(dev, dev + bytes, bytes, b200::nccl::ncclUint8, comm->comm, ctx->stream);
    if (r != b200::nccl::ncclSuccess) {
        cudaFree(dev);
        b200::set_error("ncclAllGather (set-up): %s", a.GetErrorString(r));
        return B200_ERR_COMM;
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(all, dev + bytes, bytes * comm->nranks, cudaMemcpyDeviceToHost,
                                    ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(dev);
    return B200_OK;
}

// unmap the peers' blocks; the own block is either freed (set-up failure: nobody uses it
// any more, all ranks fail alike) or handed to `retire`
void window_close(b200_peer_window* w, int rank, std::vector<void*>* retire = nullptr)
{
    for (size_t p = 0; p < w->mapped.size(); ++p)
        if ((int)p != rank && w->mapped[p]) cudaIpcCloseMemHandle(w->mapped[p]);
    w->mapped.clear();
    if (w->local) {
        if (retire)
            retire->push_back(w->local);
        else
            cudaFree(w->local);
    }
    w->local = nullptr;
}

// Collective: allocate `bytes` (zeroed), exchange IPC handles, map every peer's block.
// Every rank learns whether ALL ranks succeeded; on any failure nothing stays mapped.
b200_status window_open(b200_ctx* ctx, b200_comm* comm, size_t bytes, b200_peer_window* w)
{
    const int n = comm->nranks;
    struct Msg {
        cudaIpcMemHandle_t handle;
        int32_t ok;
        int32_t pad[3];
    };
    static_assert(sizeof(Msg) % 16 == 0, "message layout");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    w->bytes = bytes;
    mine.ok = cudaMalloc(&w->local, bytes) == cudaSuccess &&
              cudaMemsetAsync(w->local, 0, bytes, ctx->stream) == cudaSuccess &&
              cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.handle, w->local) == cudaSuccess;
    if (!mine.ok) cudaGetLastError();
    std::vector<Msg> all(n);
    b200_status st = allgather_host(ctx, comm, &mine, sizeof(Msg), all.data());
    if (st != B200_OK) return st;
    bool ok = true;
    for (int p = 0; p < n; ++p) ok = ok && all[p].ok;
    int32_t opened = ok;
    w->mapped.assign(n, nullptr);
    if (ok) {
        for (int p = 0; p < n && opened; ++p) {
            if (p == comm->rank) {
                w->mapped[p] = w->local;
                continue;
            }
            cudaError_t e = cudaIpcOpenMemHandle(&w->mapped[p], all[p].handle,
                                                 cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                b200::set_error("cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
                cudaGetLastError();
                w->mapped[p] = nullptr;
                opened = 0;
            }
        }
    }
    std::vector<int32_t> flags(n);
    st = allgather_host(ctx, comm, &opened, sizeof(int32_t), flags.data());
    if (st != B200_OK) return st;
    for (int p = 0; p < n; ++p) ok = ok && flags[p];
    if (!ok) {
        window_close(w, comm->rank, &comm->retired);
        if (!mine.ok || opened) b200::set_error("peer window: a rank could not export or map its block");
        return B200_ERR_COMM;
    }
    return B200_OK;
}

}  // namespace

extern "C" {

// Collective over the communicator: switch the <= 8-value all-reduces to the peer-memory
// kernel.  Returns B200_ERR_COMM (on every rank alike) if CUDA IPC is not available between
// the processes; the NCCL path then stays in place.
b200_status b200_comm_enable_p2p(b200_ctx* ctx, b200_comm* comm)
{
    B200_REQUIRE(ctx && comm, "null argument");
    if (comm->p2p || comm->nranks == 1) return B200_OK;
    B200_REQUIRE(comm->nranks <= 64, "peer-memory all-reduce supports up to 64 ranks");
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    comm->device = ctx->device;
    comm->ctx = ctx;
    const size_t bytes =
        2 * (size_t)comm->nranks * 8 + 2 * (size_t)comm->nranks * b200::dist::kMailCells * 8;
    b200_status st = window_open(ctx, comm, bytes, &comm->win);
    if (st != B200_OK) return st;
    if (!comm->err_dev) {
        B200_CUDA_CHECK(cudaMalloc((void**)&comm->err_dev, sizeof(int)));
        B200_CUDA_CHECK(cudaMemsetAsync(comm->err_dev, 0, sizeof(int), ctx->stream));
    }
    B200_CUDA_CHECK(cudaMalloc((void**)&comm->peer_base_dev, sizeof(void*) * comm->nranks));
    B200_CUDA_CHECK(cudaMalloc((void**)&comm->epoch_dev, sizeof(uint64_t)));
    B200_CUDA_CHECK(cudaMemsetAsync(comm->epoch_dev, 0, sizeof(uint64_t), ctx->stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(comm->peer_base_dev, comm->win.mapped.data(),
                                    sizeof(void*) * comm->nranks, cudaMemcpyHostToDevice,
                                    ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    comm->p2p = true;
    return B200_OK;
}
int32_t b200_comm_p2p_enabled(const b200_comm* comm) { return comm && comm->p2p; }
/* 1 once a peer-memory wait timed out (a rank died or the call sequences diverged);  */
/* synchronises the context's stream                                                   */
int32_t b200_comm_p2p_error(b200_ctx* ctx, const b200_comm* comm)
{
    if (!ctx || !comm || !comm->err_dev) return 0;
    int v = 0;
    if (cudaMemcpyAsync(&v, comm->err_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) !=
            cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return 1;
    return v;
}
int32_t b200_halo_p2p_enabled(const b200_halo* h) { return h && h->p2p; }

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
    if (comm->ctx || !comm->retired.empty()) {
        cudaSetDevice(comm->device);
        cudaDeviceSynchronize();
        window_close(&comm->win, comm->rank, &comm->retired);
        // barrier: every rank has unmapped its peers before anybody frees an exported block
        if (comm->ctx && comm->comm && comm->err_dev &&
            b200::nccl::api().AllReduce(comm->err_dev, comm->err_dev, 1, b200::nccl::ncclInt32,
                                        b200::nccl::ncclSum, comm->comm,
                                        comm->ctx->stream) == b200::nccl::ncclSuccess)
            cudaStreamSynchronize(comm->ctx->stream);
        else
            cudaDeviceSynchronize();
        for (void* blk : comm->retired) cudaFree(blk);
        cudaFree(comm->peer_base_dev);
        cudaFree(comm->epoch_dev);
    }
    if (comm->err_dev) cudaFree(comm->err_dev);
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
    if (h->p2p) {
        cudaDeviceSynchronize();
        window_close(&h->win, h->rank, h->comm ? &h->comm->retired : nullptr);
        cudaFree(h->meta_dev);
        cudaFree(h->peer_slot_dev);
        cudaFree(h->peer_flag_dev);
        cudaFree(h->epoch_dev);
        cudaFree(h->ticket_dev);
        cudaFree(h->run_base_dev);
        cudaFree(h->dst_ticket_dev);
        if (h->push_stream) cudaStreamDestroy(h->push_stream);
        if (h->ev_ready) cudaEventDestroy(h->ev_ready);
        if (h->ev_pushed) cudaEventDestroy(h->ev_pushed);
    }
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
        if (comm->p2p && count <= b200::dist::kMailCells) {                                    \
            b200::dist::CommDev c{comm->nranks, comm->rank, comm->peer_base_dev,               \
                                  comm->epoch_dev, comm->err_dev};                             \
            b200::dist::p2p_allreduce_kernel<VT><<<1, 64, 0, ctx->stream>>>(c, buf, (int)count);\
            B200_LAUNCH_CHECK(ctx);                                                            \
            return B200_OK;                                                                    \
        }                                                                                      \
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
        if (h->p2p) {                                                                          \
            cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;                      \
            cudaStreamIsCapturing(ctx->stream, &cap_st);                                       \
            if (cap_st != cudaStreamCaptureStatusNone)                                         \
                h->captured = true;                                                            \
            else                                                                               \
                h->host_epoch++;                                                               \
            b200::dist::HaloDev d{h->nranks, h->rank, h->n_local, h->n_ghost, h->n_send,       \
                                  h->send_idx, h->meta_dev, h->peer_slot_dev,                  \
                                  h->peer_flag_dev, (uint64_t*)h->win.local,                   \
                                  {(char*)h->win.local + h->slot_off[0],                       \
                                   (char*)h->win.local + h->slot_off[1]},                      \
                                  h->epoch_dev, h->ticket_dev, h->err_dev};                    \
            const int64_t cap = (int64_t)ctx->num_sms * 4;                                     \
            if (h->runs) {                                                                     \
                b200::dist::halo_push_runs_kernel<VT><<<2 * ctx->num_sms, 256, 0, ctx->stream>>>( \
                    d, x_ext, h->run_base_dev, nullptr);                                       \
            } else {                                                                           \
                int64_t gs = b200::ceildiv(h->n_send, 256 * 4);                                \
                gs = gs < 1 ? 1 : (gs > cap ? cap : gs);                                       \
                b200::dist::halo_push_kernel<VT><<<(unsigned)gs, 256, 0, ctx->stream>>>(d, x_ext); \
            }                                                                                  \
            B200_LAUNCH_CHECK(ctx);                                                            \
            int64_t gr = b200::ceildiv(h->n_ghost, 256 * 4);                                   \
            gr = gr < 1 ? 1 : (gr > cap ? cap : gr);                                           \
            b200::dist::halo_wait_kernel<VT><<<(unsigned)gr, 256, 0, ctx->stream>>>(d, x_ext); \
            B200_LAUNCH_CHECK(ctx);                                                            \
            return B200_OK;                                                                    \
        }                                                                                      \
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
    }                                                                                          \
    /* The exchange without the landing -> ghost copy: the owned entries x_owned[0 .. n_local) go  */ \
    /* to the peers AND into this rank's own extended vector inside the peer window; *x_ext_out is */ \
    /* that vector [owned | ghosts], valid until the exchange after the next one.  Peer memory     */ \
    /* only, not capturable (the buffer alternates with the epoch): returns B200_ERR_UNSUPPORTED   */ \
    /* otherwise and the caller uses b200_halo_exchange_*.                                         */ \
    b200_status b200_halo_exchange_inplace_##V(b200_ctx* ctx, b200_comm* comm, b200_halo* h,   \
                                               const VT* x_owned, VT** x_ext_out)              \
    {                                                                                          \
        (void)comm;                                                                            \
        B200_REQUIRE(ctx && h && x_owned && x_ext_out, "null argument");                       \
        cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;                          \
        cudaStreamIsCapturing(ctx->stream, &cap_st);                                           \
        if (!h->p2p || h->captured || cap_st != cudaStreamCaptureStatusNone) {                 \
            b200::set_error("in-place halo exchange needs the peer-memory path outside graphs"); \
            return B200_ERR_UNSUPPORTED;                                                       \
        }                                                                                      \
        const int par = (int)((h->host_epoch + 1) & 1);                                        \
        VT* own = (VT*)((char*)h->win.local + h->buf_off[par]);                                \
        b200::dist::HaloDev d{h->nranks, h->rank, h->n_local, h->n_ghost, h->n_send,           \
                              h->send_idx, h->meta_dev, h->peer_slot_dev,                      \
                              h->peer_flag_dev, (uint64_t*)h->win.local,                       \
                              {(char*)h->win.local + h->slot_off[0],                           \
                               (char*)h->win.local + h->slot_off[1]},                          \
                              h->epoch_dev, h->ticket_dev, h->err_dev};                        \
        if (h->runs) {                                                                         \
            b200::dist::halo_push_runs_kernel<VT><<<2 * ctx->num_sms, 256, 0, ctx->stream>>>(  \
                d, x_owned, h->run_base_dev, own);                                             \
            B200_LAUNCH_CHECK(ctx);                                                            \
        } else {                                                                               \
            B200_CUDA_CHECK(cudaMemcpyAsync(own, x_owned, (size_t)h->n_local * sizeof(VT),     \
                                            cudaMemcpyDeviceToDevice, ctx->stream));           \
            const int64_t cap = (int64_t)ctx->num_sms * 4;                                     \
            int64_t gs = b200::ceildiv(h->n_send, 256 * 4);                                    \
            gs = gs < 1 ? 1 : (gs > cap ? cap : gs);                                           \
            b200::dist::halo_push_kernel<VT><<<(unsigned)gs, 256, 0, ctx->stream>>>(d, x_owned); \
            B200_LAUNCH_CHECK(ctx);                                                            \
        }                                                                                      \
        b200::dist::halo_wait_inplace_kernel<<<1, 256, 0, ctx->stream>>>(d);                   \
        B200_LAUNCH_CHECK(ctx);                                                                \
        h->host_epoch++;                                                                       \
        *x_ext_out = own;                                                                      \
        return B200_OK;                                                                        \
    }                                                                                          \
    /* Pipelined exchange for matrices whose ghosts are contiguous runs of every peer (dense      */ \
    /* ghosts): begin() copies the owned entries into the rank's extended vector, starts the ring- */ \
    /* ordered push on a second stream and returns the extended vector, the per-source arrival    */ \
    /* flags (device, index = source rank) and the epoch to wait for; the caller gathers from the  */ \
    /* owner blocks as they arrive (b200_csr_spmv_part_* with wait_flag = flags + src) and calls    */ \
    /* end() after the last block.  B200_ERR_UNSUPPORTED when the halo does not qualify.           */ \
    b200_status b200_halo_exchange_staged_begin_##V(b200_ctx* ctx, b200_comm* comm, b200_halo* h, \
                                                    const VT* x_owned, VT** x_ext_out,         \
                                                    const uint64_t** flags_out,                \
                                                    uint64_t* epoch_out)                       \
    {                                                                                          \
        (void)comm;                                                                            \
        B200_REQUIRE(ctx && h && x_owned && x_ext_out && flags_out && epoch_out, "null argument"); \
        cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;                          \
        cudaStreamIsCapturing(ctx->stream, &cap_st);                                           \
        if (!h->p2p || !h->runs || h->captured || cap_st != cudaStreamCaptureStatusNone) {     \
            b200::set_error("staged halo exchange needs peer memory, contiguous runs, no graphs"); \
            return B200_ERR_UNSUPPORTED;                                                       \
        }                                                                                      \
        if (!h->push_stream) {                                                                 \
            B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->push_stream, cudaStreamNonBlocking)); \
            B200_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));   \
            B200_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_pushed, cudaEventDisableTiming));  \
            B200_CUDA_CHECK(cudaMalloc((void**)&h->dst_ticket_dev, sizeof(unsigned) * h->nranks)); \
            B200_CUDA_CHECK(cudaMemset(h->dst_ticket_dev, 0, sizeof(unsigned) * h->nranks));   \
        }                                                                                      \
        const uint64_t e = h->host_epoch + 1;                                                  \
        const int par = (int)(e & 1);                                                          \
        VT* own = (VT*)((char*)h->win.local + h->buf_off[par]);                                \
        b200::dist::HaloDev d{h->nranks, h->rank, h->n_local, h->n_ghost, h->n_send,           \
                              h->send_idx, h->meta_dev, h->peer_slot_dev,                      \
                              h->peer_flag_dev, (uint64_t*)h->win.local,                       \
                              {(char*)h->win.local + h->slot_off[0],                           \
                               (char*)h->win.local + h->slot_off[1]},                          \
                              h->epoch_dev, h->ticket_dev, h->err_dev};                        \
        B200_CUDA_CHECK(cudaEventRecord(h->ev_ready, ctx->stream));                            \
        B200_CUDA_CHECK(cudaStreamWaitEvent(h->push_stream, h->ev_ready, 0));                  \
        b200::dist::halo_push_staged_kernel<VT><<<96, 256, 0, h->push_stream>>>(               \
            d, x_owned, h->run_base_dev, h->dst_ticket_dev);                                   \
        B200_LAUNCH_CHECK(ctx);                                                                \
        B200_CUDA_CHECK(cudaEventRecord(h->ev_pushed, h->push_stream));                        \
        B200_CUDA_CHECK(cudaMemcpyAsync(own, x_owned, (size_t)h->n_local * sizeof(VT),         \
                                        cudaMemcpyDeviceToDevice, ctx->stream));               \
        *x_ext_out = own;                                                                      \
        *flags_out = (const uint64_t*)h->win.local + (size_t)par * h->nranks;                  \
        *epoch_out = e;                                                                        \
        return B200_OK;                                                                        \
    }
B200_DEF_COMM(f64, double)
B200_DEF_COMM(f32, float)

/* after the last owner block has been consumed: advance the epoch on the compute stream and make
 * that stream wait for the push (it reads the caller's x_owned) */
b200_status b200_halo_exchange_staged_end(b200_ctx* ctx, b200_halo* h)
{
    B200_REQUIRE(ctx && h && h->push_stream, "no staged exchange in flight");
    b200::dist::HaloDev d{h->nranks, h->rank, h->n_local, h->n_ghost, h->n_send,
                          h->send_idx, h->meta_dev, h->peer_slot_dev,
                          h->peer_flag_dev, (uint64_t*)h->win.local,
                          {(char*)h->win.local + h->slot_off[0], (char*)h->win.local + h->slot_off[1]},
                          h->epoch_dev, h->ticket_dev, h->err_dev};
    b200::dist::halo_epoch_advance_kernel<<<1, 1, 0, ctx->stream>>>(d);
    B200_LAUNCH_CHECK(ctx);
    B200_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, h->ev_pushed, 0));
    h->host_epoch++;
    return B200_OK;
}
/* counts of the halo (host arrays of nranks entries): what this rank receives from / sends to each peer */
b200_status b200_halo_counts(const b200_halo* h, int64_t* recv_counts, int64_t* send_counts)
{
    B200_REQUIRE(h != nullptr, "null halo");
    for (int p = 0; p < h->nranks; ++p) {
        if (recv_counts) recv_counts[p] = h->recv_count[p];
        if (send_counts) send_counts[p] = h->send_count[p];
    }
    return B200_OK;
}

// all-gather of `bytes_per_rank` raw device bytes per rank (set-up exchanges of index lists
// and counts: distributed::Matrix::read_distributed)
b200_status b200_comm_allgather_bytes(b200_ctx* ctx, b200_comm* comm, const void* send, void* recv,
                                      int64_t bytes_per_rank)
{
    B200_REQUIRE(ctx && comm && (bytes_per_rank == 0 || (send && recv)), "null argument");
    B200_REQUIRE(bytes_per_rank >= 0, "negative size");
    if (bytes_per_rank == 0) return B200_OK;
    auto& a = b200::nccl::api();
    B200_NCCL_CHECK(a.AllGather(send, recv, (size_t)bytes_per_rank, b200::nccl::ncclUint8, comm->comm,
                                ctx->stream));
    return B200_OK;
}

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


// Collective over the communicator: move this halo plan onto peer memory.  Every rank
// learns from the all-gathered send counts where its segment starts in each peer's landing
// slot and how large each peer's window is.
b200_status b200_halo_enable_p2p(b200_ctx* ctx, b200_comm* comm, b200_halo* h)
{
    B200_REQUIRE(ctx && comm && h, "null argument");
    if (h->p2p || h->nranks == 1) return B200_OK;
    B200_REQUIRE(h->nranks == comm->nranks, "halo and communicator differ in size");
    B200_REQUIRE(h->nranks <= 256, "peer-memory halo supports up to 256 ranks");
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int n = h->nranks, me = comm->rank;
    h->rank = me;
    h->comm = comm;
    comm->device = ctx->device;
    comm->ctx = ctx;
    if (!comm->err_dev) {
        B200_CUDA_CHECK(cudaMalloc((void**)&comm->err_dev, sizeof(int)));
        B200_CUDA_CHECK(cudaMemsetAsync(comm->err_dev, 0, sizeof(int), ctx->stream));
    }
    h->err_dev = comm->err_dev;
    // S[q][p] = number of entries rank q sends to rank p
    std::vector<int64_t> S((size_t)n * n);
    b200_status st = allgather_host(ctx, comm, h->send_count.data(), sizeof(int64_t) * n, S.data());
    if (st != B200_OK) return st;
    int32_t consistent = 1;
    for (int q = 0; q < n; ++q)
        if (S[(size_t)q * n + me] != h->recv_count[q]) consistent = 0;
    std::vector<int32_t> cons(n);
    st = allgather_host(ctx, comm, &consistent, sizeof(int32_t), cons.data());
    if (st != B200_OK) return st;
    for (int p = 0; p < n; ++p)
        if (!cons[p]) {
            b200::set_error("halo: send and receive counts of the ranks do not match");
            return B200_ERR_INVALID;
        }
    // Two landing slots per halo are safe only when every exchanging pair of ranks sends in BOTH
    // directions (a rank can then start epoch e+2 only after its peers finished reading epoch e).
    // A one-directional coupling stays on the NCCL exchange; S is the same on every rank, so all
    // ranks take this exit together, before the collective window set-up.
    for (int q = 0; q < n; ++q)
        for (int p = q + 1; p < n; ++p)
            if ((S[(size_t)q * n + p] > 0) != (S[(size_t)p * n + q] > 0)) {
                b200::set_error("halo: ranks %d and %d exchange in one direction only; keeping the NCCL "
                                "exchange for this matrix", q, p);
                return B200_ERR_COMM;
            }
    const size_t flag_bytes = round256(2 * (size_t)n * sizeof(uint64_t));
    auto ghosts_of = [&](int p) {
        int64_t g = 0;
        for (int q = 0; q < n; ++q) g += S[(size_t)q * n + p];
        return g;
    };
    // every rank's number of owned entries: the landing slot of rank p is the ghost tail of its
    // extended vectors [n_local_p owned | ghosts_p]
    std::vector<int64_t> NL(n);
    st = allgather_host(ctx, comm, &h->n_local, sizeof(int64_t), NL.data());
    if (st != B200_OK) return st;
    auto buf_bytes_of = [&](int p) { return round256((size_t)(NL[p] + ghosts_of(p)) * h->elem + 16); };
    h->buf_off[0] = flag_bytes;
    h->buf_off[1] = flag_bytes + buf_bytes_of(me);
    h->slot_off[0] = h->buf_off[0] + (size_t)NL[me] * h->elem;
    h->slot_off[1] = h->buf_off[1] + (size_t)NL[me] * h->elem;
    st = window_open(ctx, comm, flag_bytes + 2 * buf_bytes_of(me), &h->win);
    if (st != B200_OK) return st;
    std::vector<void*> peer_slot(2 * (size_t)n);
    std::vector<uint64_t*> peer_flag(n);
    for (int p = 0; p < n; ++p) {
        int64_t off = 0;  // where my segment starts in p's ghost numbering
        for (int q = 0; q < me; ++q) off += S[(size_t)q * n + p];
        char* base = (char*)h->win.mapped[p];
        peer_slot[p] = base + flag_bytes + (size_t)(NL[p] + off) * h->elem;
        peer_slot[n + p] = base + flag_bytes + buf_bytes_of(p) + (size_t)(NL[p] + off) * h->elem;
        peer_flag[p] = (uint64_t*)base + me;
    }
    // contiguous runs?  (host check of the send list, once)
    {
        std::vector<int32_t> idx((size_t)h->n_send);
        if (h->n_send > 0)
            B200_CUDA_CHECK(cudaMemcpy(idx.data(), h->send_idx, sizeof(int32_t) * (size_t)h->n_send,
                                       cudaMemcpyDeviceToHost));
        std::vector<int64_t> base(n, 0);
        bool runs = h->n_send > 0;
        for (int p = 0; p < n && runs; ++p) {
            const int64_t o = h->send_off[p], c = h->send_count[p];
            if (c == 0) continue;
            base[p] = idx[(size_t)o];
            for (int64_t k = 1; k < c; ++k)
                if (idx[(size_t)(o + k)] != idx[(size_t)o] + k) {
                    runs = false;
                    break;
                }
        }
        if (const char* env = getenv("B200_HALO_RUNS"))
            if (atoi(env) == 0) runs = false;
        if (runs) {
            B200_CUDA_CHECK(cudaMalloc((void**)&h->run_base_dev, sizeof(int64_t) * n));
            B200_CUDA_CHECK(cudaMemcpy(h->run_base_dev, base.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice));
        }
        h->runs = runs;
    }
    std::vector<int64_t> meta(3 * (size_t)n + 1);
    for (int p = 0; p < n; ++p) {
        meta[p] = h->send_off[p];
        meta[n + 1 + p] = h->send_count[p];
        meta[2 * n + 1 + p] = h->recv_count[p];
    }
    meta[n] = h->n_send;
    B200_CUDA_CHECK(cudaMalloc((void**)&h->meta_dev, sizeof(int64_t) * meta.size()));
    B200_CUDA_CHECK(cudaMalloc((void**)&h->peer_slot_dev, sizeof(void*) * peer_slot.size()));
    B200_CUDA_CHECK(cudaMalloc((void**)&h->peer_flag_dev, sizeof(uint64_t*) * n));
    B200_CUDA_CHECK(cudaMalloc((void**)&h->epoch_dev, sizeof(uint64_t)));
    B200_CUDA_CHECK(cudaMalloc((void**)&h->ticket_dev, 2 * sizeof(unsigned)));
    B200_CUDA_CHECK(cudaMemsetAsync(h->epoch_dev, 0, sizeof(uint64_t), ctx->stream));
    B200_CUDA_CHECK(cudaMemsetAsync(h->ticket_dev, 0, 2 * sizeof(unsigned), ctx->stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h->meta_dev, meta.data(), sizeof(int64_t) * meta.size(),
                                    cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h->peer_slot_dev, peer_slot.data(),
                                    sizeof(void*) * peer_slot.size(), cudaMemcpyHostToDevice,
                                    ctx->stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(h->peer_flag_dev, peer_flag.data(), sizeof(uint64_t*) * n,
                                    cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    h->p2p = true;
    return B200_OK;
}

}  // extern "C"
