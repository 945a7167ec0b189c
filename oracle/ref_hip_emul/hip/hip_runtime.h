/* HIP-on-CPU emulation header — ORACLE infrastructure only.
 *
 * Lets g++ compile the reference's src/rocjpeg_hip_kernels.cpp unmodified for
 * the host, so the reference's own colour-conversion / layout kernels can be
 * executed here as the pin for oracle/jpeg_oracle.c's output stage.
 * Kernel launches (<<<...>>>, rewritten to EMUL_LAUNCH by oracle/build.py's
 * sed pipe) become nested loops over the grid that set the hipBlockIdx_* /
 * hipThreadIdx_* variables and call the kernel body once per thread.
 *
 * The only arithmetic this header DEFINES (rather than inherits from the
 * reference source) is __builtin_amdgcn_cvt_pk_u8_f32, whose rounding mode the
 * reference never states: round-to-nearest-even with saturation is used — the
 * same convention as the oracle and the CUDA kernels (cvt.rni.sat.u8.f32).
 */
#pragma once
#include <stdint.h>
#include <sys/types.h>
#include <cmath>
#include <cstring>

#define __global__ static
#define __device__ static
#define __host__
#define __forceinline__ inline __attribute__((always_inline))

typedef void *hipStream_t;
typedef int hipError_t;

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

static thread_local unsigned hipBlockDim_x, hipBlockDim_y, hipBlockIdx_x, hipBlockIdx_y, hipThreadIdx_x, hipThreadIdx_y;

static inline uint32_t __builtin_amdgcn_cvt_pk_u8_f32(float f, uint32_t byte_index, uint32_t old) {
    uint32_t v;
    if (!(f > 0.0f)) v = 0;
    else if (f >= 255.0f) v = 255;
    else v = (uint32_t)lrintf(f); /* nearest-even */
    uint32_t sh = 8u * (byte_index & 3u);
    return (old & ~(0xFFu << sh)) | (v << sh);
}

#define EMUL_LAUNCH(kernel, grid, block, shmem, stream)                                   \
    [&](auto... emul_args) {                                                              \
        dim3 emul_g = (grid), emul_b = (block);                                           \
        hipBlockDim_x = emul_b.x;                                                         \
        hipBlockDim_y = emul_b.y;                                                         \
        for (unsigned by = 0; by < emul_g.y; by++)                                        \
            for (unsigned bx = 0; bx < emul_g.x; bx++)                                    \
                for (unsigned ty = 0; ty < emul_b.y; ty++)                                \
                    for (unsigned tx = 0; tx < emul_b.x; tx++) {                          \
                        hipBlockIdx_x = bx; hipBlockIdx_y = by;                           \
                        hipThreadIdx_x = tx; hipThreadIdx_y = ty;                         \
                        kernel(emul_args...);                                             \
                    }                                                                     \
    }
