// sa_fused.cu -- one SA scale in ONE kernel: gather + concat + (conv+BN+ReLU) x L + max-pool + mask, tcgen05.
//
// Fuses lib/utils/layers_util.py:157-180 of the reference (idx masking, 2x group_point, centre subtraction, concat,
// the conv2d/bias/BN/ReLU stack, reduce_max, mask multiply -- ~25 TF kernels per scale, each streaming [B,M,K,C]
// through HBM) for the layers whose weights fit in shared memory (3DSSD layer1/layer2: 93% of the grouped rows).
// Nothing but the neighbour indices, the gathered source rows (L2-resident) and the pooled [B,M,C3] result touches
// global memory; activations go TMEM -> registers -> shared memory (already split in bf16 hi/lo and already in the
// 128B-swizzled K-major layout the next layer's UMMA descriptor expects).
//
// Per CTA (128 threads, thread t <-> row t of a 128-row tile = 128/K neighbour groups):
//   weights of all layers (hi+lo images, pre-swizzled on the host) are bulk-copied to shared memory once;
//   loop over tiles: gather+split -> [fence.proxy.async] -> layer-0 MMAs (one thread) -> epilogue (tcgen05.ld,
//   scale/shift/ReLU, split) -> next layer's A operand in shared memory -> ... -> last layer: redux.sync max-pool.
// Several CTAs per SM (layer1: ~30 KiB each) overlap each other's gather / MMA / epilogue phases.
//
// Precision: same bf16 hi/lo split and 3-MMA scheme as mlp_tc.cu.
#include <cuda_bf16.h>

#include "common.cuh"
#include "pool.cuh"

namespace ssd3d {

constexpr int SF_THREADS = 128;
constexpr int SF_MAX_LAYERS = 3;

struct SfParams {
    int n, c, m, ns;                 // points per scene, feature channels, queries per scene, neighbours per query
    long rows;                       // b*m*ns
    int tiles;
    const float *xyz, *points, *new_xyz;
    const int *idx, *cnt;
    int nl;
    int kp[SF_MAX_LAYERS];           // K of layer l padded to 16
    int npad[SF_MAX_LAYERS];         // N of layer l padded to 16
    int nout[SF_MAX_LAYERS];         // true N
    uint32_t w_off[SF_MAX_LAYERS];   // byte offset of the layer's weight image in the packed blob / shared memory
    uint32_t w_half[SF_MAX_LAYERS];  // bytes of the hi half (lo follows)
    uint32_t ss_off[SF_MAX_LAYERS];  // float offset of the layer's scale (shift follows at +npad) in the ss blob
    uint32_t w_total, ss_total;      // bytes / floats
    const uint8_t *w_blob;
    const float *ss_blob;
    int sspad[SF_MAX_LAYERS];        // npad rounded up to 32 (scale/shift arrays are zero padded to this)
    int rb[SF_MAX_LAYERS];           // bytes per operand row of layer l's A buffer: 32 / 64 (one block) or 128 (64-wide k-blocks)
    uint32_t bufx_bytes, bufy_bytes;
    uint32_t tmem_cols;
    float *out_f32; int ld_f32;
    __nv_bfloat16 *out_hi, *out_lo; int ld_split;
};

__device__ __forceinline__ void sf_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sf_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sf_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major operand descriptor; rb = bytes per row = swizzle span (128 -> SWIZZLE_128B, 64 -> _64B, 32 -> _32B),
// 8-row groups are 8*rb bytes apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t sf_desc(uint32_t smem_addr, int rb)
{
    const uint64_t layout = rb == 128 ? 2ull : (rb == 64 ? 4ull : 6ull);
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * rb) >> 4) << 32) | (1ull << 46) |
           (layout << 61);
}
__device__ __forceinline__ void sf_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void sf_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sf_tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sf_split_pair(float x0, float x1, uint32_t &hw, uint32_t &lw)
{
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
    hw = *reinterpret_cast<const uint32_t *>(&h);
    lw = *reinterpret_cast<const uint32_t *>(&l);
}
// byte offset of the 16-byte chunk holding columns [8*c16g, 8*c16g+8) of row r inside an A buffer laid out as
// k-blocks of { hi tile | lo tile } (128 rows x rb bytes each), each tile in the canonical K-major swizzled layout
// of span rb: chunk j of row r sits at chunk j ^ ((r >> log2(128/rb)) & (rb/16 - 1)).
__device__ __forceinline__ uint32_t sf_a_chunk(int r, int c16g, int rb)
{
    const int nc = rb >> 4;                                  // chunks per row: 8 / 4 / 2
    const int sh = rb == 128 ? 0 : (rb == 64 ? 1 : 2);
    const int kb = c16g / nc, c16 = c16g % nc;
    return (uint32_t)kb * (uint32_t)(2 * 128 * rb) + (uint32_t)(r >> 3) * (uint32_t)(8 * rb) + (uint32_t)(r & 7) * (uint32_t)rb +
           (uint32_t)((c16 ^ ((r >> sh) & (nc - 1))) << 4);
}
__device__ __forceinline__ uint32_t sf_f2ord(float x)
{
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float sf_ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

template <int POOL>
__device__ __forceinline__ void sf_pool_store(const SfParams &p, float (&v)[32], int lane, int q, int tile,
                                              int col0, int nout, float *xs)
{
    constexpr int GP = POOL >= 32 ? 32 : POOL;
    constexpr int KEEP = 32 / GP;
    constexpr int WPG = POOL > 32 ? POOL / 32 : 1;
    warp_colmax_transpose<GP>(v, lane);                   // lane owns columns (lane % GP) * KEEP + k in v[k]
    if (POOL > 32) {                                      // groups spanning 2 or 4 warps: combine through shared memory
        xs[q * 32 + lane] = v[0];
        __syncthreads();
        if ((q % WPG) == 0) {
#pragma unroll
            for (int w = 1; w < WPG; w++) v[0] = fmaxf(v[0], xs[(q + w) * 32 + lane]);
        }
        __syncthreads();
        if ((q % WPG) != 0) return;
    }
    const int gg = tile * (128 / POOL) + (q * 32 + lane) / POOL;
    if ((long)gg * POOL >= p.rows) return;
    const bool masked = p.cnt && p.cnt[gg] == 0;
#pragma unroll
    for (int k = 0; k < KEEP; k++) {
        const int col = col0 + (lane % GP) * KEEP + k;
        if (col >= nout) continue;
        const float mx = masked ? 0.0f : v[k];
        if (p.out_f32) p.out_f32[(size_t)gg * p.ld_f32 + col] = mx;
        if (p.out_hi) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(mx);
            p.out_hi[(size_t)gg * p.ld_split + col] = hb;
            p.out_lo[(size_t)gg * p.ld_split + col] = __float2bfloat16_rn(mx - __bfloat162float(hb));
        }
    }
}

__global__ void __launch_bounds__(SF_THREADS, 6)
sa_fused_kernel(const SfParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *bufx = smem;                                // A operand of layers 0 and 2
    uint8_t *bufy = bufx + p.bufx_bytes;                 // A operand of layer 1
    uint8_t *wsm = bufy + p.bufy_bytes;                  // weight images (1024-aligned: buffers are multiples of 32 KiB)
    float *ss = reinterpret_cast<float *>(wsm + p.w_total);

    __shared__ unsigned long long w_bar, mma_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float xs[4 * 32];

    const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;

    if (tid == 0) {
        mbar_init(smem_u32(&w_bar), 1);
        mbar_init(smem_u32(&mma_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(smem_u32(&w_bar), p.w_total);
        bulk_g2s(smem_u32(wsm), p.w_blob, p.w_total, smem_u32(&w_bar));
    }
    if (q == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = tid; i < p.ss_total; i += SF_THREADS) ss[i] = __ldg(p.ss_blob + i);
    sf_fence_before();
    __syncthreads();
    sf_fence_after();
    const uint32_t tmem = tmem_base_smem;
    mbar_wait_cta(smem_u32(&w_bar), 0);                  // weights landed (async proxy writes, read by UMMA only)

    uint32_t mma_phase = 0;
    const int r = tid;                                   // row of the tile this thread owns
    const int k0 = p.c + 3;

    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        // ---- gather + centre-subtract + concat + split -> bufx (layers_util.py:157-165)
        {
            const uint32_t row = (uint32_t)tile * 128u + (uint32_t)r;  // rows < 2^31 (checked by the launcher)
            const bool ok = (long)row < p.rows;
            const uint32_t qi = ok ? row / (uint32_t)p.ns : 0u;        // == scene*m + query
            const uint32_t scene = qi / (uint32_t)p.m;
            const int a = ok ? __ldg(p.idx + row) : 0;
            const float *src_f = p.points + ((size_t)scene * p.n + a) * p.c;
            const float *src_x = p.xyz + ((size_t)scene * p.n + a) * 3;
            const float *ctr = p.new_xyz + (size_t)qi * 3;
            const int nchunk = p.kp[0] >> 3;
            for (int cg = 0; cg < nchunk; cg++) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int k = cg * 8 + e;
                    float val = 0.0f;
                    if (ok) {
                        if (k < p.c) val = __ldg(src_f + k);
                        else if (k < k0) val = __ldg(src_x + (k - p.c)) - __ldg(ctr + (k - p.c));
                    }
                    f[e] = val;
                }
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int t = 0; t < 4; t++) sf_split_pair(f[2 * t], f[2 * t + 1], hw[t], lw[t]);
                const uint32_t off = sf_a_chunk(r, cg, p.rb[0]);
                *reinterpret_cast<uint4 *>(bufx + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(bufx + off + 128 * p.rb[0]) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
        for (int l = 0; l < p.nl; l++) {
            uint8_t *ain = (l & 1) ? bufy : bufx;
            uint8_t *aout = (l & 1) ? bufx : bufy;
            // operand written with ordinary stores -> make it visible to the tensor-core (async) proxy
            sf_fence_async_smem();
            sf_fence_before();
            __syncthreads();
            if (tid == 0) {
                sf_fence_after();
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npad[l] >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint32_t abase = smem_u32(ain);
                const uint32_t wbase = smem_u32(wsm + p.w_off[l]);
                const uint32_t wtile = (uint32_t)p.npad[l] * (uint32_t)p.rb[l];
                const int nks = p.kp[l] >> 4;
                const int rb = p.rb[l];
                const uint32_t atile = 128u * (uint32_t)rb;
                for (int ks = 0; ks < nks; ks++) {
                    const int kb = ks >> 2, kin = ks & 3;            // A uses 64-wide k-blocks only when rb == 128 (else kb == 0)
                    const uint64_t a_hi = sf_desc(abase + kb * (2 * atile) + kin * 32, rb);
                    const uint64_t a_lo = sf_desc(abase + kb * (2 * atile) + atile + kin * 32, rb);
                    const uint64_t b_hi = sf_desc(wbase + kb * wtile + kin * 32, rb);     // weights use the same row span
                    const uint64_t b_lo = sf_desc(wbase + p.w_half[l] + kb * wtile + kin * 32, rb);
                    sf_mma(tmem, a_hi, b_hi, idesc, ks ? 1u : 0u);
                    sf_mma(tmem, a_lo, b_hi, idesc, 1u);
                    sf_mma(tmem, a_hi, b_lo, idesc, 1u);
                }
                sf_commit(smem_u32(&mma_bar));
            }
            mbar_wait_cta(smem_u32(&mma_bar), mma_phase);
            mma_phase ^= 1u;
            sf_fence_after();
            // ---- epilogue of layer l
            const float *sc = ss + p.ss_off[l];
            const float *sh = sc + p.sspad[l];
            const bool last = l == p.nl - 1;
            const int nchunks = (p.npad[l] + 31) / 32;
            for (int ci = 0; ci < nchunks; ci++) {
                const int c0 = ci * 32;
                uint32_t rr[32];
                sf_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, rr);
                float v[32];
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {                  // scale/shift are zero-padded to 32-column chunks
                    const float4 s4 = *reinterpret_cast<const float4 *>(sc + c0 + j4);
                    const float4 h4 = *reinterpret_cast<const float4 *>(sh + c0 + j4);
                    v[j4 + 0] = fmaxf(fmaf(__uint_as_float(rr[j4 + 0]), s4.x, h4.x), 0.0f);   // every conv has a ReLU
                    v[j4 + 1] = fmaxf(fmaf(__uint_as_float(rr[j4 + 1]), s4.y, h4.y), 0.0f);
                    v[j4 + 2] = fmaxf(fmaf(__uint_as_float(rr[j4 + 2]), s4.z, h4.z), 0.0f);
                    v[j4 + 3] = fmaxf(fmaf(__uint_as_float(rr[j4 + 3]), s4.w, h4.w), 0.0f);
                }
                if (!last) {
#pragma unroll
                    for (int j8 = 0; j8 < 32; j8 += 8) {
                        if (c0 + j8 >= p.kp[l + 1]) break;
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int t = 0; t < 4; t++) sf_split_pair(v[j8 + 2 * t], v[j8 + 2 * t + 1], hw[t], lw[t]);
                        const uint32_t off = sf_a_chunk(r, (c0 + j8) >> 3, p.rb[l + 1]);
                        *reinterpret_cast<uint4 *>(aout + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4 *>(aout + off + 128 * p.rb[l + 1]) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                } else {
                    switch (p.ns) {
                        case 8: sf_pool_store<8>(p, v, lane, q, tile, c0, p.nout[l], xs); break;
                        case 16: sf_pool_store<16>(p, v, lane, q, tile, c0, p.nout[l], xs); break;
                        case 32: sf_pool_store<32>(p, v, lane, q, tile, c0, p.nout[l], xs); break;
                        case 64: sf_pool_store<64>(p, v, lane, q, tile, c0, p.nout[l], xs); break;
                        default: sf_pool_store<128>(p, v, lane, q, tile, c0, p.nout[l], xs); break;
                    }
                }
            }
            sf_fence_before();                                         // TMEM reads done before the next MMAs overwrite it
        }
        __syncthreads();                                               // bufx free for the next tile's gather
    }

    sf_fence_before();
    __syncthreads();
    if (q == 0) {
        sf_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace ssd3d

using namespace ssd3d;

// A-operand buffer of a layer with padded K = kp: { hi | lo } tiles of 128 rows; rows are 32 / 64 bytes when the
// whole K fits (kp 16 / 32), else 128-byte rows in 64-wide k-blocks.  Rounded up to 1 KiB (descriptor alignment).
static int sf_row_bytes(int kp) { return kp <= 16 ? 32 : (kp <= 32 ? 64 : 128); }
static size_t sf_abuf_bytes(int kp)
{
    const int rb = sf_row_bytes(kp);
    const int nkb = rb == 128 ? (kp + 63) / 64 : 1;
    return ((size_t)nkb * 2 * 128 * rb + 1023) / 1024 * 1024;
}

// Shared-memory bytes of the fused kernel for a layer stack, or 0 when it cannot hold it.
extern "C" size_t ssd3d_sa_fused_smem(int c, int nl, const int *nout)
{
    if (nl < 1 || nl > SF_MAX_LAYERS) return 0;
    int kp[SF_MAX_LAYERS], npad[SF_MAX_LAYERS];
    size_t w = 0, ssf = 0;
    int kprev = (c + 3 + 15) / 16 * 16;
    for (int l = 0; l < nl; l++) {
        kp[l] = kprev;
        npad[l] = (nout[l] + 15) / 16 * 16;
        if (npad[l] > 256) return 0;
        const int rbw = sf_row_bytes(kp[l]);
        const int nkb = rbw == 128 ? (kp[l] + 63) / 64 : 1;
        w += (size_t)2 * nkb * npad[l] * rbw;
        ssf += (size_t)2 * ((npad[l] + 31) / 32 * 32);
        kprev = npad[l];
    }
    const size_t a0 = sf_abuf_bytes(kp[0]), a1 = nl > 1 ? sf_abuf_bytes(kp[1]) : 0, a2 = nl > 2 ? sf_abuf_bytes(kp[2]) : 0;
    const size_t bufx = a0 > a2 ? a0 : a2;
    const size_t bufy = a1;
    const size_t total = bufx + bufy + w + ssf * sizeof(float) + 1024;
    return total <= 226 * 1024 ? total : 0;
}

// w_blob: per layer { hi image | lo image }, each image = k-blocks of [npad x 64] bf16 in the canonical
// K-major SWIZZLE_128B layout (built by params.FusedStack); ss_blob: per layer { scale[npad] | shift[npad] }.
extern "C" int ssd3d_sa_mlp_fused(int b, int n, int c, int m, int nsample, const float *xyz, const float *points,
                                  const float *new_xyz, const int *idx, const int *pts_cnt, int nl, const int *nout,
                                  const void *w_blob, const float *ss_blob, float *out_f32, int ld_f32, void *out_hi,
                                  void *out_lo, int ld_split, ssd3d_stream_t stream)
{
    SSD3D_REQUIRE(b >= 0 && n > 0 && m >= 0 && c >= 0, "sa_mlp_fused: bad shape");
    SSD3D_REQUIRE(nsample == 8 || nsample == 16 || nsample == 32 || nsample == 64 || nsample == 128,
                  "sa_mlp_fused: nsample=%d must be one of 8, 16, 32, 64, 128", nsample);
    SSD3D_REQUIRE(xyz && new_xyz && idx && w_blob && ss_blob && (points || c == 0), "sa_mlp_fused: null pointer");
    SSD3D_REQUIRE(out_f32 || (out_hi && out_lo), "sa_mlp_fused: no output requested");
    SSD3D_REQUIRE((reinterpret_cast<uintptr_t>(w_blob) & 15u) == 0, "sa_mlp_fused: weight blob must be 16-byte aligned");
    const size_t smem = ssd3d_sa_fused_smem(c, nl, nout);
    if (smem == 0) { set_error("sa_mlp_fused: layer stack does not fit shared memory"); return SSD3D_ERR_UNSUPPORTED; }
    SfParams p = {};
    p.n = n; p.c = c; p.m = m; p.ns = nsample;
    p.rows = (long)b * m * nsample;
    if (p.rows == 0) return 0;
    SSD3D_REQUIRE(p.rows < (1L << 31) - 128, "sa_mlp_fused: too many grouped rows (%ld)", p.rows);
    p.tiles = (int)((p.rows + 127) / 128);
    p.xyz = xyz; p.points = points; p.new_xyz = new_xyz; p.idx = idx; p.cnt = pts_cnt;
    p.nl = nl;
    int kprev = (c + 3 + 15) / 16 * 16;
    uint32_t woff = 0, ssoff = 0;
    int maxn = 32;
    for (int l = 0; l < nl; l++) {
        p.kp[l] = kprev;
        p.nout[l] = nout[l];
        p.npad[l] = (nout[l] + 15) / 16 * 16;
        const int rbw = sf_row_bytes(p.kp[l]);
        const int nkb = rbw == 128 ? (p.kp[l] + 63) / 64 : 1;
        p.w_off[l] = woff;
        p.w_half[l] = (uint32_t)nkb * p.npad[l] * rbw;
        woff += 2 * p.w_half[l];
        p.ss_off[l] = ssoff;
        p.sspad[l] = (p.npad[l] + 31) / 32 * 32;
        ssoff += 2 * p.sspad[l];
        kprev = p.npad[l];
        if (p.npad[l] > maxn) maxn = p.npad[l];
    }
    p.w_total = woff; p.ss_total = ssoff;
    p.w_blob = (const uint8_t *)w_blob; p.ss_blob = ss_blob;
    for (int l = 0; l < nl; l++) p.rb[l] = sf_row_bytes(p.kp[l]);
    const size_t a0 = sf_abuf_bytes(p.kp[0]), a1 = nl > 1 ? sf_abuf_bytes(p.kp[1]) : 0, a2 = nl > 2 ? sf_abuf_bytes(p.kp[2]) : 0;
    p.bufx_bytes = (uint32_t)(a0 > a2 ? a0 : a2);
    p.bufy_bytes = (uint32_t)a1;
    uint32_t cols = 32;
    while ((int)cols < maxn) cols *= 2;
    p.tmem_cols = cols;
    p.out_f32 = out_f32; p.ld_f32 = ld_f32;
    p.out_hi = (__nv_bfloat16 *)out_hi; p.out_lo = (__nv_bfloat16 *)out_lo; p.ld_split = ld_split;

    cudaError_t e = cudaFuncSetAttribute((const void *)sa_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "sa_mlp_fused attr");
    // resident CTAs per SM: bounded by shared memory and by TMEM columns (512 per SM)
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > (int)(512 / cols)) per_sm = 512 / cols;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;                                       // register file: 6 x 128 threads x 80 registers
    int grid = kNumSMs * per_sm;
    if (grid > p.tiles) grid = p.tiles;
    sa_fused_kernel<<<grid, SF_THREADS, smem, (cudaStream_t)stream>>>(p);
    SSD3D_LAUNCH_CHECK("sa_fused_kernel");
}
