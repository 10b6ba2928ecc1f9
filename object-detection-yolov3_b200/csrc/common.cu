// common.cu - error plumbing and buffer helpers.
#include "common.cuh"
#include <stdarg.h>

namespace y3 {

void fail(y3_status code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error{code, std::string(buf)};
}

void DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return;
    release();
    bytes = (bytes + 255) & ~size_t(255);
    Y3_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
void PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return;
    release();
    Y3_CUDA(cudaMallocHost(&p, bytes));
    cap = bytes;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}
void EventTimer::init() {
    if (!a) { Y3_CUDA(cudaEventCreate(&a)); Y3_CUDA(cudaEventCreate(&b)); }
}
void EventTimer::destroy() {
    if (a) { cudaEventDestroy(a); cudaEventDestroy(b); a = b = nullptr; }
}
float EventTimer::ms() {
    float t = 0.f;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&t, a, b);
    return t;
}

}  // namespace y3
