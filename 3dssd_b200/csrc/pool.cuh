// pool.cuh -- max-pool over the rows of a warp-held tile: "reduce and transpose" butterfly.
#pragma once

namespace ssd3d {

// Each lane holds one ROW of a [32 rows x 32 columns] block in v[0..31].  Computes, for every run of GP consecutive
// lanes (rows), the column-wise maximum; on return lane L owns columns (L % GP) * (32/GP) + k, k < 32/GP, of its
// group's result in v[k].  31 shuffles for GP = 32 (instead of 32 full-warp reductions): every step halves the
// number of live values per lane while exchanging the other half with the xor-partner.
template <int GP>
__device__ __forceinline__ void warp_colmax_transpose(float (&v)[32], int lane)
{
#pragma unroll
    for (int o = GP / 2, cnt = 32; o >= 1; o >>= 1, cnt >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; i++) {
            const float send = up ? v[i] : v[i + cnt / 2];
            const float keep = up ? v[i + cnt / 2] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
        }
    }
}

}  // namespace ssd3d
