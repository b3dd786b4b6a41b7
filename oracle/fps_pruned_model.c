/* fps_pruned_model.c -- TEST INFRASTRUCTURE ONLY (like everything under oracle/).
 *
 * A sequential, lane-by-lane CPU model of the PRODUCT's pruned D-FPS kernel (3dssd_b200/csrc/fps_bucket.cu), so that the
 * claim that kernel rests on -- "skipping a bucket whose box is farther from the new sample than its largest running
 * distance never changes an index, and the (value, key-with-place) arg-max returns the reference's winner" -- is
 * checked on the CPU against the restatement of the reference (oracle_farthest_point_sample, i.e.
 * /root/reference/lib/utils/tf_ops/sampling/tf_sampling_g.cu:124-178) without a GPU.  The model follows the kernel's
 * data structures one for one: 2-D Morton keys of the two widest axes (9 bits each, key = morton18:index14), buckets of
 * 32 in the interleaved order fb_pos(warp, slot, lane), a box + cached (max bits, key) per bucket, the skip rule
 *     max == 0  ||  (lb * 0.99999f >= max  &&  lb >= 1e-30f),
 * the per-warp cached best with its dirty flag, the 16-candidate scene arg-max, and the resume state in `temp`
 * (running distances in ORIGINAL order + the bucket permutation).  The box distance lb is written as
 * gx*gx + gy*gy + gz*gz in the kernel, which nvcc is free to contract; `contract` selects the fma chain (1) or three
 * separately rounded products (0) -- the exactness argument must hold for both, and the tests run both.
 *
 * Returns per-launch statistics through `stats` (may be NULL): [0] = bucket updates summed over rounds,
 * [1] = rounds run, [2] = the largest number of bucket updates in one round.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FB_NW 16
#define FB_SLOTS 32
#define FB_NBUCKET (FB_NW * FB_SLOTS)
#define FB_MAXN (FB_NBUCKET * 32)
#define KEY_INVALID 0x7FFFFFFFu
#define FB_PAD_KEY 0xFFFFFFFFu

static uint32_t f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static float u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }

static uint32_t spread9(uint32_t v)
{
    v &= 0x1ffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
static int fb_pos(int w, int i, int lane) { return ((i * FB_NW + w) << 5) + lane; }
static uint32_t fb_key(uint32_t o, uint32_t p) { return ((o & 1023u) << 21) | ((o >> 10) << 17) | p; }
static int fb_key_to_k(uint32_t key) { return (int)((((key >> 17) & 15u) << 10) | (key >> 21)); }

static int cmp_u32(const void *a, const void *b)
{
    const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* (value, key) arg-max update: max value, then min key among equal values (warp_argmax of csrc/fps.cuh) */
static void amax(uint32_t u, uint32_t key, uint32_t *mx, uint32_t *kmin)
{
    if (u > *mx) { *mx = u; *kmin = key; }
    else if (u == *mx && key < *kmin) *kmin = key;
}

typedef struct {
    float bminx, bminy, bminz, bmaxx, bmaxy, bmaxz;
    uint32_t bmaxu, bkey;
} bucket_t;

/* one scene; data stride 3 floats per point; returns 0, or -1 for unsupported sizes */
static int pruned_scene(int n, int m, const float *data, int *idxs, int ioff, int j0, int j1, float *tsave, int contract,
                        long long *stats)
{
    if (n > FB_MAXN || n < 1 || m < 1) return -1;
    const int jbeg = j0 > 1 ? j0 : 1, jend = j1 < m ? j1 : m;
    const int resume = j0 > 0, save = j1 < m;
    uint32_t *perm = tsave ? (uint32_t *)(tsave + n) : NULL;
    if ((resume || save) && !tsave) return -1;

    uint16_t *orig_of = (uint16_t *)malloc(sizeof(uint16_t) * FB_MAXN);
    float *xs = (float *)malloc(sizeof(float) * FB_MAXN * 4), *ys = xs + FB_MAXN, *zs = ys + FB_MAXN, *dist = zs + FB_MAXN;
    bucket_t *bk = (bucket_t *)malloc(sizeof(bucket_t) * FB_NBUCKET);

    if (!resume) {
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f}, ext[3];
        for (int k = 0; k < n; k++)
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], data[3 * k + a]); hi[a] = fmaxf(hi[a], data[3 * k + a]); }
        for (int a = 0; a < 3; a++) ext[a] = hi[a] - lo[a];
        int a0 = 0, a1;
        if (ext[1] > ext[a0]) a0 = 1;
        if (ext[2] > ext[a0]) a0 = 2;
        a1 = a0 == 0 ? 1 : 0;
        for (int a = 0; a < 3; a++)
            if (a != a0 && ext[a] > ext[a1]) a1 = a;
        const float scale = ext[a0] > 0.0f ? 511.999f / ext[a0] : 0.0f;
        uint32_t *skey = (uint32_t *)malloc(sizeof(uint32_t) * FB_MAXN);
        for (int k = 0; k < n; k++) {
            const float u = (data[3 * k + a0] - lo[a0]) * scale, v = (data[3 * k + a1] - lo[a1]) * scale;
            int qu = (int)u, qv = (int)v;
            qu = qu < 0 ? 0 : (qu > 511 ? 511 : qu);
            qv = qv < 0 ? 0 : (qv > 511 ? 511 : qv);
            skey[k] = ((spread9((uint32_t)qu) | (spread9((uint32_t)qv) << 1)) << 14) | (uint32_t)k;
        }
        qsort(skey, (size_t)n, sizeof(uint32_t), cmp_u32);          /* keys are unique: any sort gives the kernel's order */
        for (int p = 0; p < FB_MAXN; p++) orig_of[p] = p < n ? (uint16_t)(skey[p] & 0x3fffu) : (uint16_t)0xffffu;
        if (tsave)
            for (int p = 0; p < n; p++) perm[p] = orig_of[p];
        free(skey);
    } else {
        for (int p = 0; p < FB_MAXN; p++) orig_of[p] = p < n ? (uint16_t)(perm[p] & 0xffffu) : (uint16_t)0xffffu;
    }

    for (int w = 0; w < FB_NW; w++)
        for (int i = 0; i < FB_SLOTS; i++) {
            bucket_t *b = &bk[w * FB_SLOTS + i];
            int any = 0;
            uint32_t mx = 0u, kmin = KEY_INVALID;
            for (int lane = 0; lane < 32; lane++) {
                const int p = fb_pos(w, i, lane);
                const uint32_t o = orig_of[p];
                const int valid = o != 0xffffu;
                float x = 0.0f, y = 0.0f, z = 0.0f;
                if (valid) { x = data[3 * o]; y = data[3 * o + 1]; z = data[3 * o + 2]; }
                xs[p] = x; ys[p] = y; zs[p] = z;
                dist[p] = valid ? (resume ? tsave[o] : 1e38f) : -1.0f;
                if (valid) {
                    if (!any) { b->bminx = b->bmaxx = x; b->bminy = b->bmaxy = y; b->bminz = b->bmaxz = z; any = 1; }
                    b->bminx = fminf(b->bminx, x); b->bmaxx = fmaxf(b->bmaxx, x);
                    b->bminy = fminf(b->bminy, y); b->bmaxy = fmaxf(b->bmaxy, y);
                    b->bminz = fminf(b->bminz, z); b->bmaxz = fmaxf(b->bmaxz, z);
                }
                amax(valid ? f2u(fmaxf(dist[p], 0.0f)) : 0u, valid ? fb_key(o, (uint32_t)p) : KEY_INVALID, &mx, &kmin);
            }
            if (!any) b->bminx = b->bminy = b->bminz = b->bmaxx = b->bmaxy = b->bmaxz = NAN;   /* as the kernel's unord(~0) */
            b->bmaxu = mx; b->bkey = kmin;
        }

    int old0 = 0;
    if (resume) old0 = idxs[jbeg - 1] - ioff;
    else idxs[0] = ioff;
    float sx = data[3 * old0], sy = data[3 * old0 + 1], sz = data[3 * old0 + 2];
    uint32_t wm[FB_NW], wk[FB_NW];
    int wdirty[FB_NW];
    for (int w = 0; w < FB_NW; w++) { wm[w] = 0u; wk[w] = KEY_INVALID; wdirty[w] = 1; }
    long long total = 0, worst = 0;

    for (int j = jbeg; j < jend; j++) {
        long long touched = 0;
        for (int w = 0; w < FB_NW; w++) {
            for (int i = 0; i < FB_SLOTS; i++) {
                bucket_t *b = &bk[w * FB_SLOTS + i];
                const float gx = fmaxf(fmaxf(b->bminx - sx, sx - b->bmaxx), 0.0f);
                const float gy = fmaxf(fmaxf(b->bminy - sy, sy - b->bmaxy), 0.0f);
                const float gz = fmaxf(fmaxf(b->bminz - sz, sz - b->bmaxz), 0.0f);
                float lb;
                if (contract) lb = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
                else { const float a = gx * gx, c = gy * gy, e = gz * gz; lb = (a + c) + e; }
                const int skip = b->bmaxu == 0u || (lb * 0.99999f >= u2f(b->bmaxu) && lb >= 1e-30f);
                if (skip) continue;
                wdirty[w] = 1;
                touched++;
                uint32_t mx = 0u, kmin = KEY_INVALID;
                for (int lane = 0; lane < 32; lane++) {
                    const int p = fb_pos(w, i, lane);
                    const uint32_t kk = orig_of[p];
                    const float dx = xs[p] - sx, dy = ys[p] - sy, dz = zs[p] - sz;
                    float d = dx * dx;
                    d = fmaf(dy, dy, d);
                    d = fmaf(dz, dz, d);
                    const float nd = fminf(d, dist[p]);
                    dist[p] = nd;
                    const int valid = kk != 0xffffu;
                    amax(valid ? f2u(fmaxf(nd, 0.0f)) : 0u, valid ? fb_key(kk, (uint32_t)p) : KEY_INVALID, &mx, &kmin);
                }
                b->bmaxu = mx; b->bkey = kmin;
            }
            if (wdirty[w]) {
                uint32_t mx = 0u, kmin = KEY_INVALID;
                for (int i = 0; i < FB_SLOTS; i++) amax(bk[w * FB_SLOTS + i].bmaxu, bk[w * FB_SLOTS + i].bkey, &mx, &kmin);
                wm[w] = mx; wk[w] = kmin;
                wdirty[w] = 0;
            }
        }
        uint32_t m3 = 0u, k3 = KEY_INVALID;
        for (int w = 0; w < FB_NW; w++) amax(wm[w], wk[w], &m3, &k3);
        const int p = (int)(k3 & 0x3fffu);
        sx = xs[p]; sy = ys[p]; sz = zs[p];
        idxs[j] = fb_key_to_k(k3) + ioff;
        total += touched;
        if (touched > worst) worst = touched;
    }

    if (save)
        for (int p = 0; p < FB_MAXN; p++)
            if (orig_of[p] != 0xffffu) tsave[orig_of[p]] = dist[p];
    if (stats) { stats[0] += total; stats[1] += jend > jbeg ? jend - jbeg : 0; if (worst > stats[2]) stats[2] = worst; }
    free(orig_of); free(xs); free(bk);
    return 0;
}

/* inp [b,n,3] dense, out [b,ldo] (index j of scene s at out[s*ldo + j], value + ioff), temp [b,2n] or NULL */
int oracle_fps_pruned_model(int b, int n, int m, const float *inp, int *out, int ldo, int ioff, int j0, int j1, float *temp,
                            int contract, long long *stats)
{
    if (stats) stats[0] = stats[1] = stats[2] = 0;
    for (int s = 0; s < b; s++) {
        const int rc = pruned_scene(n, m, inp + (size_t)s * n * 3, out + (size_t)s * ldo, ioff, j0, j1,
                                    temp ? temp + (size_t)s * 2 * n : NULL, contract, stats);
        if (rc) return rc;
    }
    return 0;
}
