/* hip/hip_runtime.h — minimal HIP-on-CUDA shim.
 *
 * The reference's public header includes "hip/hip_runtime.h"
 * (api/rocjpeg.h:28) and its samples call a handful of HIP host functions
 * (samples/rocjpeg_samples_utils.h:55-61,244-264,580-610;
 * samples/jpegDecode/jpegdecode.cpp:136-196). This file maps exactly that
 * surface onto the CUDA runtime so those sources build unmodified with
 * g++/nvcc against librocjpeg.so. It is NOT a general HIP layer: nothing
 * inside the decoder uses it.
 */
#ifndef ROCJPEG_B200_HIP_RUNTIME_SHIM_H
#define ROCJPEG_B200_HIP_RUNTIME_SHIM_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#else
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#endif

typedef cudaError_t hipError_t;
typedef cudaStream_t hipStream_t;
#define hipSuccess cudaSuccess

/* Only the fields the samples print (rocjpeg_samples_utils.h:259-261). */
typedef struct hipDeviceProp_t {
    char name[256];
    char gcnArchName[256];
    int pciBusID;
    int pciDomainID;
    int pciDeviceID;
    int multiProcessorCount;
    size_t totalGlobalMem;
} hipDeviceProp_t;

static inline hipError_t hipGetDeviceCount(int *count) { return cudaGetDeviceCount(count); }
static inline hipError_t hipSetDevice(int device_id) { return cudaSetDevice(device_id); }
static inline hipError_t hipFree(void *ptr) { return cudaFree(ptr); }
static inline hipError_t hipMemcpyDtoH(void *dst, const void *src, size_t size_bytes) {
    return cudaMemcpy(dst, src, size_bytes, cudaMemcpyDeviceToHost);
}
static inline hipError_t hipMemcpyHtoD(void *dst, const void *src, size_t size_bytes) {
    return cudaMemcpy(dst, src, size_bytes, cudaMemcpyHostToDevice);
}
static inline hipError_t hipDeviceSynchronize(void) { return cudaDeviceSynchronize(); }
static inline const char *hipGetErrorName(hipError_t e) { return cudaGetErrorName(e); }
static inline const char *hipGetErrorString(hipError_t e) { return cudaGetErrorString(e); }

static inline hipError_t hipGetDeviceProperties(hipDeviceProp_t *prop, int device_id) {
    struct cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device_id);
    if (e != cudaSuccess) return e;
    memset(prop, 0, sizeof(*prop));
    strncpy(prop->name, p.name, sizeof(prop->name) - 1);
    snprintf(prop->gcnArchName, sizeof(prop->gcnArchName), "sm_%d%d%s", p.major, p.minor,
             (p.major >= 9) ? "a" : "");
    prop->pciBusID = p.pciBusID;
    prop->pciDomainID = p.pciDomainID;
    prop->pciDeviceID = p.pciDeviceID;
    prop->multiProcessorCount = p.multiProcessorCount;
    prop->totalGlobalMem = p.totalGlobalMem;
    return cudaSuccess;
}

#ifdef __cplusplus
/* The samples pass uint8_t** (jpegdecode.cpp:154). */
template <typename T>
static inline hipError_t hipMalloc(T **ptr, size_t size_bytes) {
    return cudaMalloc(reinterpret_cast<void **>(ptr), size_bytes);
}
#else
static inline hipError_t hipMalloc(void **ptr, size_t size_bytes) { return cudaMalloc(ptr, size_bytes); }
#endif

#endif /* ROCJPEG_B200_HIP_RUNTIME_SHIM_H */
