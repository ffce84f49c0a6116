// backend.h -- the few runtime calls the C-ABI layer needs (allocate, copy, stream).
// The shipped library is always built by nvcc against the CUDA runtime.  Defining
// IBIS_B200_EMU (done ONLY by tests/emu/Makefile) maps the same calls to host memory so
// that the ABI layer and the kernel bodies can be exercised in the GPU-less development
// container; that build is a test harness with a different file name and is never
// loaded by the ibis_b200 package.
#pragma once

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef IBIS_B200_EMU
#include <cuda_runtime.h>
#endif

namespace ib {
namespace gpu {

#ifndef IBIS_B200_EMU
typedef cudaStream_t stream_t;
inline int set_device(int dev) { return dev >= 0 ? (int)cudaSetDevice(dev) : 0; }
inline int get_device(int* dev) { return (int)cudaGetDevice(dev); }
inline int stream_create(stream_t* s) { return (int)cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
inline int stream_destroy(stream_t s) { return (int)cudaStreamDestroy(s); }
inline int stream_sync(stream_t s) { return (int)cudaStreamSynchronize(s); }
inline int dev_malloc(void** p, size_t bytes) { return (int)cudaMalloc(p, bytes ? bytes : 8); }
inline int dev_free(void* p) { return (int)cudaFree(p); }
inline int host_malloc(void** p, size_t bytes) { return (int)cudaMallocHost(p, bytes); }
inline int host_free(void* p) { return (int)cudaFreeHost(p); }
inline int memset_async(void* p, int v, size_t bytes, stream_t s) { return (int)cudaMemsetAsync(p, v, bytes, s); }
inline int h2d(void* d, const void* h, size_t bytes, stream_t s) {
    return (int)cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
}
inline int d2h(void* h, const void* d, size_t bytes, stream_t s) {
    return (int)cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
}
inline int d2d(void* dst, const void* src, size_t bytes, stream_t s) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
}
typedef cudaEvent_t event_t;
inline int event_create(event_t* e) { return (int)cudaEventCreateWithFlags(e, cudaEventDisableTiming); }
inline int event_destroy(event_t e) { return (int)cudaEventDestroy(e); }
inline int event_record(event_t e, stream_t s) { return (int)cudaEventRecord(e, s); }
inline int stream_wait_event(stream_t s, event_t e) { return (int)cudaStreamWaitEvent(s, e, 0); }
inline int event_sync(event_t e) { return (int)cudaEventSynchronize(e); }
inline const char* error_string(int e) { return cudaGetErrorString((cudaError_t)e); }
inline int last_error() { return (int)cudaGetLastError(); }
inline int sm_count(int dev) {
    int n = 0;
    return cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0 ? n : 148;
}
#else
typedef void* stream_t;
inline int set_device(int) { return 0; }
inline int get_device(int* dev) { *dev = 0; return 0; }
inline int stream_create(stream_t* s) { *s = nullptr; return 0; }
inline int stream_destroy(stream_t) { return 0; }
inline int stream_sync(stream_t) { return 0; }
inline int dev_malloc(void** p, size_t bytes) { *p = calloc(bytes ? bytes : 8, 1); return *p ? 0 : 2; }
inline int dev_free(void* p) { free(p); return 0; }
inline int host_malloc(void** p, size_t bytes) { *p = calloc(bytes, 1); return *p ? 0 : 2; }
inline int host_free(void* p) { free(p); return 0; }
inline int memset_async(void* p, int v, size_t bytes, stream_t) { memset(p, v, bytes); return 0; }
inline int h2d(void* d, const void* h, size_t bytes, stream_t) { memcpy(d, h, bytes); return 0; }
inline int d2h(void* h, const void* d, size_t bytes, stream_t) { memcpy(h, d, bytes); return 0; }
inline int d2d(void* dst, const void* src, size_t bytes, stream_t) { memmove(dst, src, bytes); return 0; }
typedef void* event_t;
inline int event_create(event_t* e) { *e = nullptr; return 0; }
inline int event_destroy(event_t) { return 0; }
inline int event_record(event_t, stream_t) { return 0; }
inline int stream_wait_event(stream_t, event_t) { return 0; }
inline int event_sync(event_t) { return 0; }
inline const char* error_string(int) { return "host emulation error"; }
inline int last_error() { return 0; }
inline int sm_count(int) { return 1; }
#endif

}  // namespace gpu
}  // namespace ib
