// Host-side launch helpers shared by the kernel translation units: per-device facts and one-time
// kernel attributes with thread-safe lazy initialisation (the C ABI promises re-entrancy: several
// host threads, several devices per process), and a small cache of encoded TMA tensor maps.
#pragma once
#include <atomic>
#include <cuda.h>

#include "common.cuh"

namespace sba {

constexpr int kMaxDevices = 64;

// Current device and its SM count (cached per device; concurrent first calls write the same value).
inline int current_device(int* dev_out, int* sms_out, const char* what) {
    static std::atomic<int> sms_of[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error("%s: device index %d not supported (%s)", what, dev, cudaGetErrorString(e));
        return SBA_ERR_UNSUPPORTED;
    }
    int n = sms_of[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || n < 1) {
            set_error("%s: cudaDeviceGetAttribute(SM count): %s", what, cudaGetErrorString(e));
            return SBA_ERR_CUDA;
        }
        sms_of[dev].store(n, std::memory_order_relaxed);
    }
    *dev_out = dev;
    *sms_out = n;
    return SBA_OK;
}

// cuTensorMapEncodeTiled is a driver entry point: it needs the device's primary context to be current on the calling
// host thread, which only a runtime call that touches the device guarantees.  A thread that arrives after the one-time
// attribute calls below have been made by another thread (e.g. autograd's backward thread after a first call from the
// main thread) has made no such call yet: bind the context once per thread (found by a test that ran in that order:
// CUDA_ERROR_INVALID_CONTEXT from the encoder).
inline void bind_context_to_thread() {
    thread_local int bound_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != bound_dev) {
        cudaFree(nullptr);
        bound_dev = dev;
    }
}

// Raise a kernel's dynamic shared-memory limit once per device.  `done` is a call-site static bit mask
// (one bit per device); setting the attribute twice from racing threads is harmless.
template <class Kern>
inline int ensure_dynamic_smem(Kern kern, size_t bytes, int dev, std::atomic<unsigned long long>& done, const char* what) {
    const unsigned long long bit = 1ull << dev;
    if (done.load(std::memory_order_acquire) & bit) return SBA_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%zu B of dynamic shared memory): %s", what, bytes, cudaGetErrorString(e));
        return SBA_ERR_CUDA;
    }
    done.fetch_or(bit, std::memory_order_release);
    return SBA_OK;
}

// launch attribute block of a programmatic dependent launch (the kernel may become resident while its
// predecessor in the stream is still running; it orders itself with griddepcontrol.wait)
struct PdlLaunch {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st) : cfg{} {
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
};

}  // namespace sba
