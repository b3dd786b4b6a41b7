// fps.cuh -- pieces shared by the FPS kernels (fps.cu, fps_bucket.cu).
#pragma once
#include "common.cuh"

namespace ssd3d {

constexpr uint32_t KEY_INVALID = 0x7FFFFFFFu;

// Where a launch reads and writes (the *_ex entry points of include/ssd3d.h): scene strides of the inputs in
// elements (so a [:, a:b] slice of a dense [b,N,c] tensor needs no copy), row stride + value offset of the index
// output (so the segments of a fusion-sampling layer land directly in the concatenated fps_idx tensor with the
// segment offset already added, layers_util.py:109-111), and -- fps3_direct_kernel only -- the range of rounds
// [j0, j1) this launch runs, the running distances travelling through `temp` between launches.
struct FpsIO {
    long long sa, sb;   // scene stride of inp / fa (sa) and fb (sb), in floats
    int ldo, ioff;      // out row stride (ints), offset added to every stored index
    int j0, j1;         // rounds [j0, j1) of 0..m (round 0 = "sample point 0")
    float *temp;        // [b, n] running distances (resume state; NULL when j0 == 0 && j1 == m)
};

__device__ __forceinline__ uint32_t fps_key(int k) { return ((uint32_t)(k & 1023) << 21) | (uint32_t)(k >> 10); }
__device__ __forceinline__ int fps_key_to_k(uint32_t key) { return (int)(((key & 0x1FFFFFu) << 10) | (key >> 21)); }

// (value,key) arg-max across a warp: max value, then min key among the lanes holding it.
__device__ __forceinline__ void warp_argmax(uint32_t u, uint32_t key_if_valid, uint32_t &mx, uint32_t &kmin)
{
    mx = __reduce_max_sync(0xffffffffu, u);
    kmin = __reduce_min_sync(0xffffffffu, (u == mx) ? key_if_valid : KEY_INVALID);
}


// fps_bucket.cu: single-CTA D-FPS with spatial pruning for large xyz scenes
bool fps3_bucket_applies(int n, int m, const float *inp, long long sa, int flags);
int launch_fps3_bucket(int b, int n, int m, const float *inp, int *out, const FpsIO &io, cudaStream_t st);

}  // namespace ssd3d
