/* bq_grid_model.c -- TEST INFRASTRUCTURE ONLY (like everything under oracle/).
 *
 * A sequential CPU model of the PRODUCT's culled ball query (3dssd_b200/csrc/ball_query_grid.cu): uniform 2-D grid over
 * the two widest axes (cell >= 1.01 * r_max, <= 8192 cells, fp32 cell arithmetic), counting sort into per-cell records,
 * a query visits the three contiguous record ranges of its 3x3 cell neighbourhood, hits are marked in a per-shell bitmap
 * over candidate indices and read back in ascending index (first nsample, first hit back-filled, cnt = min(hits, nsample));
 * neighbourhoods holding more than n/8 candidates -- and queries with a non-finite coordinate -- take the index-order
 * scan; every non-empty group lists ceil(cnt / 8) units (group << 4 | j).  It lets the CPU tests compare the kernel's
 * ALGORITHM with the restatement of the reference (oracle_query_ball_point[_dilated], i.e.
 * /root/reference/lib/utils/tf_ops/grouping/tf_grouping_g.cu:215-255, :308-357) neighbour list for neighbour list.
 * Squared-distance thresholds T(r) = smallest float t with sqrt_rn(t) >= r are computed here the way the product's host
 * code does (csrc/ball_query.cu sq_threshold), so that  sqrt(t) < r  <=>  t < T(r).
 *
 * stats (may be NULL): [0] queries that took the dense (index-order) path, [1] candidates streamed by the sparse path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BQG_MAX_N 16384
#define BQG_MAX_CELLS 8192
#define BQG_MAX_SHELLS 4

static float sq_threshold(float r)
{
    if (!(r > 0.0f)) return 0.0f;
    if (isinf(r)) return INFINITY;
    float t = (float)((double)r * (double)r);
    if (isinf(t)) t = 3.402823466e+38f;
    while (t > 0.0f && sqrtf(nextafterf(t, 0.0f)) >= r) t = nextafterf(t, 0.0f);
    while (!isinf(t) && sqrtf(t) < r) t = nextafterf(t, INFINITY);
    return t;
}

static int cell_coord(float v, float mn, float inv_c, int nc)
{
    const float u = (v - mn) * inv_c;
    int i = isnan(u) ? 0 : (u <= -2.0e9f ? -2000000000 : (u >= 2.0e9f ? 2000000000 : (int)floorf(u)));
    i = i < 0 ? 0 : i;
    return i > nc - 1 ? nc - 1 : i;
}

typedef struct { float x, y, z; int k; } rec_t;

static int hit_of(int dilated, float t, float t_lo, float t_hi)
{
    return dilated ? (t == 0.0f || (t >= t_lo && t < t_hi)) : !(t >= t_hi);
}

int oracle_bq_grid_model(int b, int n, int m, int nshell, int dilated, const float *min_radius, const float *max_radius,
                         const int *nsample, const float *xyz1, const float *xyz2, int **idx, int **cnt, int **units,
                         long long *stats)
{
    if (n <= 0 || n > BQG_MAX_N || nshell < 1 || nshell > BQG_MAX_SHELLS) return -1;
    float t_lo[BQG_MAX_SHELLS], t_hi[BQG_MAX_SHELLS], t_max = 0.0f, r_max = 0.0f;
    for (int s = 0; s < nshell; s++) {
        t_hi[s] = sq_threshold(max_radius[s]);
        t_lo[s] = dilated ? sq_threshold(min_radius[s]) : 0.0f;
        if (!dilated && !(max_radius[s] > 1e-20f)) t_hi[s] = -1.0f;
        t_max = fmaxf(t_max, t_hi[s]);
        r_max = fmaxf(r_max, max_radius[s]);
        if (units && units[s]) units[s][0] = 0;
    }
    if (stats) stats[0] = stats[1] = 0;
    int *cell_start = (int *)malloc(sizeof(int) * (BQG_MAX_CELLS + 2));
    int *fill = (int *)malloc(sizeof(int) * (BQG_MAX_CELLS + 2));
    rec_t *rec = (rec_t *)malloc(sizeof(rec_t) * (size_t)n);
    uint32_t *bm = (uint32_t *)malloc(sizeof(uint32_t) * (BQG_MAX_N / 32));

    for (int scene = 0; scene < b; scene++) {
        const float *pts = xyz1 + (size_t)scene * n * 3;
        /* ---- grid of the scene */
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, ext[3];
        int bad = 0;
        for (int k = 0; k < n; k++)
            for (int a = 0; a < 3; a++) {
                const float v = pts[3 * k + a];
                if (!(fabsf(v) <= 3.0e38f)) bad = 1;
                mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v);
            }
        for (int a = 0; a < 3; a++) ext[a] = mx[a] - mn[a];
        int a0 = 0, a1 = 1, a2 = 2, tmp;
        if (ext[a1] > ext[a0]) { tmp = a0; a0 = a1; a1 = tmp; }
        if (ext[a2] > ext[a0]) { tmp = a0; a0 = a2; a2 = tmp; }
        if (ext[a2] > ext[a1]) { tmp = a1; a1 = a2; a2 = tmp; }
        float min_a = 0.0f, min_b = 0.0f, inv_c = 0.0f;
        int na = 1, nb = 1;
        const int one_cell = bad || !(ext[a0] >= 0.0f) || !(ext[a0] <= 1.0e30f) || !(r_max > 0.0f) || !(r_max <= 1.0e30f);
        if (!one_cell) {
            float c = r_max * 1.01f;
            for (;;) {
                const float fa = floorf(ext[a0] / c) + 1.0f, fb = floorf(ext[a1] / c) + 1.0f;
                if (fa * fb <= (float)BQG_MAX_CELLS) { na = (int)fa; nb = (int)fb; break; }
                c *= 1.25f;
            }
            min_a = mn[a0]; min_b = mn[a1]; inv_c = 1.0f / c;
        }
        const int ncell = na * nb;
        memset(cell_start, 0, sizeof(int) * (BQG_MAX_CELLS + 2));
        for (int k = 0; k < n; k++)
            cell_start[1 + cell_coord(pts[3 * k + a1], min_b, inv_c, nb) * na + cell_coord(pts[3 * k + a0], min_a, inv_c, na)]++;
        for (int i = 0; i < ncell; i++) cell_start[i + 1] += cell_start[i];
        memcpy(fill, cell_start, sizeof(int) * (ncell + 1));
        for (int k = n - 1; k >= 0; k--) {      /* DESCENDING: the order inside a cell must not matter (the GPU's is arbitrary) */
            const int c = cell_coord(pts[3 * k + a1], min_b, inv_c, nb) * na + cell_coord(pts[3 * k + a0], min_a, inv_c, na);
            rec_t r = { pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], k };
            rec[fill[c]++] = r;
        }

        for (int qi = 0; qi < m; qi++) {
            const float *q = xyz2 + ((size_t)scene * m + qi) * 3;
            const float qx = q[0], qy = q[1], qz = q[2];
            const int qfin = fabsf(qx) <= 3.0e38f && fabsf(qy) <= 3.0e38f && fabsf(qz) <= 3.0e38f;
            const int ca = cell_coord(q[a0], min_a, inv_c, na), cb = cell_coord(q[a1], min_b, inv_c, nb);
            const int a_lo = ca > 0 ? ca - 1 : 0, a_hi = ca + 1 < na ? ca + 1 : na - 1;
            const int b_lo = cb > 0 ? cb - 1 : 0, b_hi = cb + 1 < nb ? cb + 1 : nb - 1;
            int j0r[3], j1r[3], ncand = 0;
            for (int rr = 0; rr < 3; rr++) {
                const int rb = b_lo + rr;
                j0r[rr] = rb <= b_hi ? cell_start[rb * na + a_lo] : 0;
                j1r[rr] = rb <= b_hi ? cell_start[rb * na + a_hi + 1] : 0;
                ncand += j1r[rr] - j0r[rr];
            }
            const size_t g = (size_t)scene * m + qi;
            if (!qfin || ncand * 8 > n) {
                if (stats) stats[0]++;
                for (int s = 0; s < nshell; s++) {
                    const int ns = nsample[s];
                    int *dst = idx[s] + g * ns, c = 0, first = 0;
                    for (int k = 0; k < n && c < ns; k++) {
                        const float dx = qx - pts[3 * k], dy = qy - pts[3 * k + 1], dz = qz - pts[3 * k + 2];
                        float t = dy * dy;
                        t = fmaf(dx, dx, t);
                        t = fmaf(dz, dz, t);
                        if (hit_of(dilated, t, t_lo[s], t_hi[s])) { if (c == 0) first = k; dst[c++] = k; }
                    }
                    for (int l = c; l < ns; l++) dst[l] = first;
                    cnt[s][g] = c;
                    if (units && units[s] && c > 0) {
                        const int nu = (c + 7) >> 3, base = units[s][0];
                        units[s][0] += nu;
                        for (int j = 0; j < nu; j++) units[s][1 + base + j] = (int)((g << 4) | (size_t)j);
                    }
                }
                continue;
            }
            if (stats) stats[1] += ncand;
            for (int s = 0; s < nshell; s++) {
                memset(bm, 0, sizeof(uint32_t) * (BQG_MAX_N / 32));
                for (int rr = 0; rr < 3; rr++)
                    for (int j = j0r[rr]; j < j1r[rr]; j++) {
                        const rec_t c = rec[j];
                        const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
                        float t = dy * dy;
                        t = fmaf(dx, dx, t);
                        t = fmaf(dz, dz, t);
                        const int near = dilated ? (t < t_max) : !(t >= t_max);
                        if (near && hit_of(dilated, t, t_lo[s], t_hi[s])) bm[c.k >> 5] |= 1u << (c.k & 31);
                    }
                const int ns = nsample[s];
                int *dst = idx[s] + g * ns, total = 0, pos = 0, first = 0;
                for (int w = 0; w < (n + 31) / 32; w++)
                    for (int bit = 0; bit < 32; bit++)
                        if (bm[w] >> bit & 1u) {
                            const int k = w * 32 + bit;
                            if (total == 0) first = k;
                            if (pos < ns) dst[pos++] = k;
                            total++;
                        }
                const int c = total < ns ? total : ns;
                for (int l = c; l < ns; l++) dst[l] = c > 0 ? first : 0;
                cnt[s][g] = c;
                if (units && units[s] && c > 0) {
                    const int nu = (c + 7) >> 3, base = units[s][0];
                    units[s][0] += nu;
                    for (int j = 0; j < nu; j++) units[s][1 + base + j] = (int)((g << 4) | (size_t)j);
                }
            }
        }
    }
    free(cell_start); free(fill); free(rec); free(bm);
    return 0;
}
