// devbuf.h -- tiny RAII device buffer + CUDA error check shared by the host-side translation units.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#define CUDA_OK(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace pg {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    // grows geometrically: buffers that creep up by a few elements per call (sample counts, work lists) must not
    // pay a synchronising cudaFree + cudaMalloc every time
    void alloc(size_t count) {
        if (count <= n) return;
        release();
        const size_t want = std::max<size_t>(count + count / 2, 64);
        CUDA_OK(cudaMalloc(&p, want * sizeof(T)));
        n = want;
    }
    void allocExact(size_t count) {  // capacity == count (buffers whose .n is used as the logical size)
        if (count == n) return;
        release();
        CUDA_OK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        n = count;
    }
    void upload(const T *src, size_t count, cudaStream_t st = 0) {
        alloc(count);
        if (count) CUDA_OK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void upload(const std::vector<T> &v, cudaStream_t st = 0) { upload(v.data(), v.size(), st); }
};

}  // namespace pg
