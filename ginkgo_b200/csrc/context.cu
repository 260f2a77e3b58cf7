// Executor glue behind include/ginkgo_b200.h: device/stream ownership, raw
// memory, copies, synchronisation.  Mirrors what CudaExecutor provides to the
// kernels in the reference (cuda/base/executor.cpp: raw_alloc/raw_free/
// raw_copy_to/synchronize; stubs listed at core/device_hooks/cuda_hooks.cpp:19-250).
#include <stdarg.h>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace b200

void* b200_ctx::scratch(size_t bytes)
{
    if (bytes <= ws.bytes) return ws.ptr;
    // grow: stream-ordered free + alloc keeps earlier kernels valid
    size_t nb = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
    void* np = nullptr;
    if (cudaMallocAsync(&np, nb, stream) != cudaSuccess) return nullptr;
    if (ws.ptr) cudaFreeAsync(ws.ptr, stream);
    ws.ptr = np;
    ws.bytes = nb;
    return np;
}

__global__ void snapshot_copy_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int bytes)
{
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}

extern "C" {

const char* b200_last_error(void) { return b200::g_err; }

const char* b200_version(void)
{
    return "ginkgo_b200 0.1 (hand-written CUDA, sm_100a, CUDA " B200_STR(CUDART_VERSION) ")";
}

b200_status b200_ctx_create(int32_t device_id, void* cuda_stream, b200_ctx** out)
{
    B200_REQUIRE(out != nullptr, "out must not be null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        b200::set_error("no CUDA device available (%s): this library has no CPU fallback",
                        cudaGetErrorString(e));
        return B200_ERR_CUDA;
    }
    B200_REQUIRE(device_id >= 0 && device_id < ndev, "device_id out of range");
    B200_CUDA_CHECK(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10) {
        b200::set_error("device %d is sm_%d%d; this library contains sm_100a code only", device_id,
                        prop.major, prop.minor);
        return B200_ERR_UNSUPPORTED;
    }
    b200_ctx* c = new b200_ctx();
    c->device = device_id;
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->owns_stream = true;
    }
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaMalloc(&c->snap_dev, 2 * b200_ctx::kSnapBytes));
    B200_CUDA_CHECK(cudaMallocHost((void**)&c->snap_host, 2 * b200_ctx::kSnapBytes));
    for (int i = 0; i < 2; ++i) B200_CUDA_CHECK(cudaEventCreateWithFlags(&c->snap_ev[i], cudaEventDisableTiming));
    {
        // keep released blocks in the device pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    B200_CUDA_CHECK(cudaMalloc((void**)&c->counters, 256 * sizeof(unsigned int)));
    B200_CUDA_CHECK(cudaMemsetAsync(c->counters, 0, 256 * sizeof(unsigned int), c->stream));
    B200_CUDA_CHECK(cudaMallocHost((void**)&c->pinned, 4096));
    memset(c->pinned, 0, 4096);
    B200_CUDA_CHECK(cudaMalloc(&c->dev_mailbox, 4096));
    B200_CUDA_CHECK(cudaMemsetAsync(c->dev_mailbox, 0, 4096, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *out = c;
    return B200_OK;
}

void b200_ctx_destroy(b200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->aux) {
        cudaStreamSynchronize(ctx->aux);
        cudaStreamDestroy(ctx->aux);
    }
    if (ctx->fork) cudaEventDestroy(ctx->fork);
    if (ctx->join) cudaEventDestroy(ctx->join);
    for (int i = 0; i < 2; ++i)
        if (ctx->snap_ev[i]) cudaEventDestroy(ctx->snap_ev[i]);
    if (ctx->snap_dev) cudaFree(ctx->snap_dev);
    if (ctx->snap_host) cudaFreeHost(ctx->snap_host);
    if (ctx->ws.ptr) cudaFree(ctx->ws.ptr);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->dev_mailbox) cudaFree(ctx->dev_mailbox);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

/* Stream-ordered snapshot: the bytes at src_dev AS THEY ARE when the stream reaches this point are
 * copied into slot `slot` (device to device, on the stream) and from there to pinned host memory on
 * the auxiliary stream; b200_snapshot_end waits for that copy only -- work enqueued on the stream
 * after the begin keeps running.  The fused solvers poll their control block this way: the next
 * batch of iterations is already queued when the host looks at the previous one, so the GPU never
 * idles for a host round trip (and the ranks of a distributed solve all see the state at the SAME
 * batch boundary, i.e. take the same decision). */
b200_status b200_snapshot_begin(b200_ctx* ctx, int32_t slot, const void* src_dev, size_t bytes)
{
    B200_REQUIRE(ctx && src_dev, "null argument");
    B200_REQUIRE(slot >= 0 && slot < 2 && bytes <= b200_ctx::kSnapBytes, "bad snapshot slot / size");
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    char* d = (char*)ctx->snap_dev + (size_t)slot * b200_ctx::kSnapBytes;
    // a 64-thread kernel, not cudaMemcpyAsync: the copy stays on the compute engine (no hand-over to a
    // copy engine and back between two graph launches)
    snapshot_copy_kernel<<<1, 64, 0, ctx->stream>>>((const unsigned char*)src_dev, (unsigned char*)d, (int)bytes);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaEventRecord(ctx->snap_ev[slot], ctx->stream));
    B200_CUDA_CHECK(cudaStreamWaitEvent(ctx->aux, ctx->snap_ev[slot], 0));
    B200_CUDA_CHECK(cudaMemcpyAsync(ctx->snap_host + (size_t)slot * b200_ctx::kSnapBytes, d, bytes,
                                    cudaMemcpyDeviceToHost, ctx->aux));
    B200_CUDA_CHECK(cudaEventRecord(ctx->snap_ev[slot], ctx->aux));
    return B200_OK;
}

b200_status b200_snapshot_end(b200_ctx* ctx, int32_t slot, void* dst_host, size_t bytes)
{
    B200_REQUIRE(ctx && dst_host, "null argument");
    B200_REQUIRE(slot >= 0 && slot < 2 && bytes <= b200_ctx::kSnapBytes, "bad snapshot slot / size");
    B200_CUDA_CHECK(cudaEventSynchronize(ctx->snap_ev[slot]));
    memcpy(dst_host, ctx->snap_host + (size_t)slot * b200_ctx::kSnapBytes, bytes);
    return B200_OK;
}

void* b200_ctx_stream(const b200_ctx* ctx) { return (void*)ctx->stream; }
int32_t b200_ctx_device(const b200_ctx* ctx) { return ctx->device; }
int32_t b200_ctx_num_sms(const b200_ctx* ctx) { return ctx->num_sms; }
int64_t b200_ctx_launch_count(const b200_ctx* ctx) { return ctx->launches; }

b200_status b200_alloc(b200_ctx* ctx, size_t bytes, void** out)
{
    B200_REQUIRE(ctx && out, "null argument");
    *out = nullptr;
    if (bytes == 0) return B200_OK;
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    // stream-ordered allocation (the reference's CudaAsyncAllocator, cuda/base/memory.cpp):
    // solver workspaces are allocated and released on every apply; with the pool they are
    // re-used without a device synchronisation or a page-mapping cost
    cudaError_t e = cudaMallocAsync(out, bytes, ctx->stream);
    if (e != cudaSuccess) {
        b200::set_error("cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return B200_ERR_ALLOC;
    }
    return B200_OK;
}

b200_status b200_free(b200_ctx* ctx, void* ptr)
{
    // like CudaExecutor::raw_free this must not fail loudly
    // (core/device_hooks/cuda_hooks.cpp:113-118)
    if (!ptr) return B200_OK;
    cudaSetDevice(ctx->device);
    if (cudaFreeAsync(ptr, ctx->stream) != cudaSuccess) cudaGetLastError();
    return B200_OK;
}

b200_status b200_copy_h2d(b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) return B200_OK;
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

b200_status b200_copy_d2h(b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) return B200_OK;
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

b200_status b200_copy_d2d(b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) return B200_OK;
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
}

b200_status b200_synchronize(b200_ctx* ctx)
{
    B200_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}


// ---------------------------------------------------------------------------------------
// Staging pipe: host-resident operands of an apply (the reference clones them onto the
// device inside LinOp::apply, include/ginkgo/core/base/lin_op.hpp:129-215 +
// make_temporary_clone) move over two extra streams, so that the upload of call k+1, the
// kernel of call k and the download of call k-1 overlap (PCIe is full duplex).  `slot`
// selects one of kPipeSlots staging-buffer pairs the caller owns; the events enforce
//   upload(slot)   after the compute that last read the slot's input buffer,
//   compute(slot)  after upload(slot) and after the download that last read its output,
//   download(slot) after compute(slot).
// ---------------------------------------------------------------------------------------
struct b200_pipe {
    static constexpr int kSlots = 2;
    b200_ctx* ctx = nullptr;
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t uploaded[kSlots], computed[kSlots], downloaded[kSlots];
};

b200_status b200_pipe_create(b200_ctx* ctx, b200_pipe** out)
{
    B200_REQUIRE(ctx && out, "null argument");
    B200_CUDA_CHECK(cudaSetDevice(ctx->device));
    b200_pipe* p = new b200_pipe();
    p->ctx = ctx;
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&p->in, cudaStreamNonBlocking));
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&p->out, cudaStreamNonBlocking));
    for (int i = 0; i < b200_pipe::kSlots; ++i) {
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&p->uploaded[i], cudaEventDisableTiming));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&p->computed[i], cudaEventDisableTiming));
        B200_CUDA_CHECK(cudaEventCreateWithFlags(&p->downloaded[i], cudaEventDisableTiming));
    }
    // Staging buffers come from the stream-ordered pool of ctx->stream: a block handed to the owner of
    // this pipe may still be in use by earlier work on ctx->stream (plan tuning, solver workspaces freed
    // with b200_free).  Order both copy streams after everything enqueued on ctx->stream so far, and
    // pre-record `computed` on ctx->stream so the first upload / download of every slot has a real
    // dependency instead of waiting on a never-recorded event (a no-op).
    cudaEvent_t born;
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&born, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventRecord(born, ctx->stream));
    B200_CUDA_CHECK(cudaStreamWaitEvent(p->in, born, 0));
    B200_CUDA_CHECK(cudaStreamWaitEvent(p->out, born, 0));
    for (int i = 0; i < b200_pipe::kSlots; ++i) B200_CUDA_CHECK(cudaEventRecord(p->computed[i], ctx->stream));
    B200_CUDA_CHECK(cudaEventDestroy(born));
    *out = p;
    return B200_OK;
}

void b200_pipe_destroy(b200_pipe* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    // the owner frees its staging buffers on ctx->stream after this call: nothing of the copy streams
    // may still touch them (host-synchronous, so the stream-ordered frees that follow are safe)
    cudaStreamSynchronize(p->in);
    cudaStreamSynchronize(p->out);
    for (int i = 0; i < b200_pipe::kSlots; ++i) {
        cudaEventDestroy(p->uploaded[i]);
        cudaEventDestroy(p->computed[i]);
        cudaEventDestroy(p->downloaded[i]);
    }
    cudaStreamDestroy(p->in);
    cudaStreamDestroy(p->out);
    delete p;
}

int32_t b200_pipe_num_slots(void) { return b200_pipe::kSlots; }

/* host -> device on the upload stream (src_host should be pinned to overlap) */
b200_status b200_pipe_upload(b200_pipe* p, int32_t slot, void* dst_dev, const void* src_host,
                             size_t bytes)
{
    B200_REQUIRE(p && slot >= 0 && slot < b200_pipe::kSlots, "bad pipe slot");
    B200_CUDA_CHECK(cudaStreamWaitEvent(p->in, p->computed[slot], 0));
    if (bytes)
        B200_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, p->in));
    B200_CUDA_CHECK(cudaEventRecord(p->uploaded[slot], p->in));
    return B200_OK;
}

/* bracket the kernels of one call on the context's stream */
b200_status b200_pipe_begin_compute(b200_pipe* p, int32_t slot)
{
    B200_REQUIRE(p && slot >= 0 && slot < b200_pipe::kSlots, "bad pipe slot");
    B200_CUDA_CHECK(cudaStreamWaitEvent(p->ctx->stream, p->uploaded[slot], 0));
    B200_CUDA_CHECK(cudaStreamWaitEvent(p->ctx->stream, p->downloaded[slot], 0));
    return B200_OK;
}
b200_status b200_pipe_end_compute(b200_pipe* p, int32_t slot)
{
    B200_REQUIRE(p && slot >= 0 && slot < b200_pipe::kSlots, "bad pipe slot");
    B200_CUDA_CHECK(cudaEventRecord(p->computed[slot], p->ctx->stream));
    return B200_OK;
}

/* device -> host on the download stream */
b200_status b200_pipe_download(b200_pipe* p, int32_t slot, void* dst_host, const void* src_dev,
                               size_t bytes)
{
    B200_REQUIRE(p && slot >= 0 && slot < b200_pipe::kSlots, "bad pipe slot");
    B200_CUDA_CHECK(cudaStreamWaitEvent(p->out, p->computed[slot], 0));
    if (bytes)
        B200_CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, p->out));
    B200_CUDA_CHECK(cudaEventRecord(p->downloaded[slot], p->out));
    return B200_OK;
}

/* the context's stream waits for every outstanding transfer (then b200_synchronize, or an
 * event recorded on that stream, covers the whole pipeline) */
b200_status b200_pipe_join(b200_pipe* p)
{
    B200_REQUIRE(p != nullptr, "null pipe");
    for (int i = 0; i < b200_pipe::kSlots; ++i) {
        B200_CUDA_CHECK(cudaStreamWaitEvent(p->ctx->stream, p->uploaded[i], 0));
        B200_CUDA_CHECK(cudaStreamWaitEvent(p->ctx->stream, p->downloaded[i], 0));
    }
    return B200_OK;
}

}  // extern "C"
