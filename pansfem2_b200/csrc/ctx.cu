// ctx.cu -- context, memory and timing entry points of the C ABI.
#include <cstdarg>
#include "common.cuh"

namespace pf2 {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
__global__ void flush_kernel(double* buf, size_t n, double v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}
}  // namespace pf2

extern "C" {

const char* pf2_last_error(void) { return pf2::g_err; }
const char* pf2_version(void) { return "pansfem2_b200 0.1 (sm_100a)"; }

int pf2_ctx_create(int device, void* stream, pf2_ctx** out) {
    PF2_CHECK(out != nullptr, "null output");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        pf2::set_error("no CUDA device available (%s): libpansfem2_b200 has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return PF2_E_NODEVICE;
    }
    PF2_CHECK(device >= 0 && device < count, "device index out of range");
    PF2_CUDA(cudaSetDevice(device));
    pf2_ctx* c = new pf2_ctx();
    c->device = device;
    cudaDeviceProp prop;
    PF2_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->total_mem = prop.totalGlobalMem;
    c->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    const char* env = getenv("PF2_L2_PERSIST");
    c->l2_persist_enabled = (env && env[0] == '1');   // opt-in: the set-aside slows the sub-warp SpMV kernels (r01 sweep)
    if (c->l2_persist_enabled && c->l2_persist_max > 0) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_max) != cudaSuccess) { cudaGetLastError(); c->l2_persist_enabled = false; }
    }
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { PF2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    PF2_TRY(pf2::dev_alloc(&c->red.partials, (size_t)pf2::kMaxBlocks * pf2::kMaxTerms));
    PF2_TRY(pf2::dev_alloc(&c->red.ticket, 4));
    PF2_CUDA(cudaMemsetAsync(c->red.ticket, 0, 4 * sizeof(unsigned int), c->stream));
    PF2_TRY(pf2::dev_alloc(&c->scalars, 64));
    PF2_CUDA(cudaMemsetAsync(c->scalars, 0, 64 * sizeof(double), c->stream));
    PF2_CUDA(cudaHostAlloc((void**)&c->h_scalars, 64 * sizeof(double), cudaHostAllocDefault));
    PF2_CUDA(cudaEventCreate(&c->ev0));
    PF2_CUDA(cudaEventCreate(&c->ev1));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return PF2_OK;
}

int pf2_ctx_destroy(pf2_ctx* c) {
    if (!c) return PF2_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->red.partials); cudaFree(c->red.ticket); cudaFree(c->scalars); cudaFreeHost(c->h_scalars);
    if (c->flush_buf) cudaFree(c->flush_buf);
    if (c->elem_scratch) cudaFree(c->elem_scratch);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return PF2_OK;
}

int pf2_ctx_sync(pf2_ctx* c) { PF2_CUDA(cudaStreamSynchronize(c->stream)); return PF2_OK; }

int pf2_ctx_device_info(pf2_ctx* c, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (total_mem) *total_mem = c->total_mem;
    return PF2_OK;
}
int pf2_ctx_launch_count(pf2_ctx* c, long long* out) { *out = c->launches; return PF2_OK; }

int pf2_malloc(pf2_ctx* c, size_t bytes, void** dev_out) {
    PF2_CUDA(cudaSetDevice(c->device));
    PF2_CUDA(cudaMalloc(dev_out, bytes ? bytes : 8));
    return PF2_OK;
}
int pf2_free(pf2_ctx* c, void* dev) {
    if (!dev) return PF2_OK;
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    PF2_CUDA(cudaFree(dev));
    return PF2_OK;
}
int pf2_memcpy_h2d(pf2_ctx* c, void* dst, const void* src, size_t bytes) {
    PF2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return PF2_OK;
}
int pf2_memcpy_d2h(pf2_ctx* c, void* dst, const void* src, size_t bytes) {
    PF2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    PF2_CUDA(cudaStreamSynchronize(c->stream));
    return PF2_OK;
}
int pf2_memset(pf2_ctx* c, void* dst, int value, size_t bytes) {
    PF2_CUDA(cudaMemsetAsync(dst, value, bytes, c->stream));
    return PF2_OK;
}
int pf2_host_alloc(size_t bytes, void** host_out) {
    PF2_CUDA(cudaHostAlloc(host_out, bytes ? bytes : 8, cudaHostAllocDefault));
    return PF2_OK;
}
int pf2_host_free(void* host) {
    if (host) PF2_CUDA(cudaFreeHost(host));
    return PF2_OK;
}
int pf2_timer_start(pf2_ctx* c) { PF2_CUDA(cudaEventRecord(c->ev0, c->stream)); return PF2_OK; }
int pf2_timer_stop(pf2_ctx* c, double* ms) {
    PF2_CUDA(cudaEventRecord(c->ev1, c->stream));
    PF2_CUDA(cudaEventSynchronize(c->ev1));
    float f = 0;
    PF2_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = f;
    return PF2_OK;
}
int pf2_flush_l2(pf2_ctx* c) {
    const size_t bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
    if (!c->flush_buf) { PF2_CUDA(cudaMalloc(&c->flush_buf, bytes)); c->flush_bytes = bytes; }
    pf2::flush_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>((double*)c->flush_buf, bytes / 8, 1.0);
    PF2_LAUNCH_CHECK();
    return PF2_OK;
}

}  // extern "C"
