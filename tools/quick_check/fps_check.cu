// fps_check.cu -- a torch-free, seconds-long GPU check of the D-FPS kernels through the C ABI: the pruned single-CTA kernel
// (flags = 4) and the cluster kernel (flags = 2) must both return the indices the CPU oracle computed for the same scenes
// (files written by tools/quick_check/make_inputs.py), incl. a run of the pruned kernel in four resumed parts; prints one
// JSON object with the verdicts and the time per round.  Exists so that a check fits into a GPU slot too short to import
// a framework.  Build: see tools/quick_check/Makefile.  Test infrastructure, not product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/ssd3d.h"

static std::vector<char> slurp(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> v(n);
    if (fread(v.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : "gpurun_in";
    char path[512];
    snprintf(path, sizeof path, "%s/fps_check_cases.bin", dir);
    std::vector<char> blob = slurp(path);
    const int *hdr = (const int *)blob.data();
    const int ncases = hdr[0];
    size_t off = 4 * (1 + 3 * (size_t)ncases);
    cudaStream_t st; cudaStreamCreate(&st);
    printf("{\"cases\": [");
    int all_ok = 1;
    for (int ci = 0; ci < ncases; ci++) {
        const int b = hdr[1 + 3 * ci], n = hdr[2 + 3 * ci], m = hdr[3 + 3 * ci];
        const float *pts = (const float *)(blob.data() + off); off += sizeof(float) * (size_t)b * n * 3;
        const int *expect = (const int *)(blob.data() + off); off += sizeof(int) * (size_t)b * m;
        float *d_pts, *d_temp; int *d_out;
        cudaMalloc(&d_pts, sizeof(float) * (size_t)b * n * 3);
        cudaMalloc(&d_temp, sizeof(float) * (size_t)b * n * 2);
        cudaMalloc(&d_out, sizeof(int) * (size_t)b * m);
        cudaMemcpy(d_pts, pts, sizeof(float) * (size_t)b * n * 3, cudaMemcpyHostToDevice);
        std::vector<int> got((size_t)b * m);
        int ok[3] = {0, 0, 0}, rc[3] = {0, 0, 0};
        float ms_bucket = 0, ms_cluster = 0;
        for (int variant = 0; variant < 3; variant++) {          // 0 pruned kernel, 1 cluster kernel, 2 pruned kernel in 4 resumed parts
            cudaMemset(d_out, 0xff, sizeof(int) * (size_t)b * m);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int flags = variant == 1 ? 2 : 4;
            if (variant < 2) {
                rc[variant] = ssd3d_farthest_point_sample_ex(b, n, 3, m, d_pts, (long long)n * 3, d_temp, d_out, m, 0, 0, m, 0, flags, st);  // warm-up
                cudaEventRecord(e0, st);
                rc[variant] |= ssd3d_farthest_point_sample_ex(b, n, 3, m, d_pts, (long long)n * 3, d_temp, d_out, m, 0, 0, m, 0, flags, st);
                cudaEventRecord(e1, st);
            } else {
                const int cuts[5] = {0, m / 3, m / 2, m - 1, m};
                cudaEventRecord(e0, st);
                for (int p = 0; p < 4; p++)
                    if (cuts[p + 1] > cuts[p])
                        rc[variant] |= ssd3d_farthest_point_sample_ex(b, n, 3, m, d_pts, (long long)n * 3, d_temp, d_out, m, 0, cuts[p], cuts[p + 1], 0, flags, st);
                cudaEventRecord(e1, st);
            }
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc[variant] |= (int)e;
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            if (variant == 0) ms_bucket = ms; else if (variant == 1) ms_cluster = ms;
            cudaMemcpy(got.data(), d_out, sizeof(int) * (size_t)b * m, cudaMemcpyDeviceToHost);
            ok[variant] = rc[variant] == 0 && memcmp(got.data(), expect, sizeof(int) * (size_t)b * m) == 0;
        }
        // round time of the pruned kernel without its per-scene setup: the first m/2 samples are a prefix of the m samples, so
        // (time of m rounds - time of m/2 rounds) / (m - m/2) is the cost of a late round; best of 5 launches each
        float best_full = 1e30f, best_half = 1e30f;
        for (int rep = 0; rep < 5 && m >= 512; rep++) {
            for (int half = 0; half < 2; half++) {
                const int mm = half ? m / 2 : m;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0, st);
                ssd3d_farthest_point_sample_ex(b, n, 3, mm, d_pts, (long long)n * 3, d_temp, d_out, m, 0, 0, mm, 0, 4, st);
                cudaEventRecord(e1, st);
                cudaStreamSynchronize(st);
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                if (half) best_half = ms < best_half ? ms : best_half; else best_full = ms < best_full ? ms : best_full;
            }
        }
        const float late_round_ns = m >= 512 ? 1e6f * (best_full - best_half) / (float)(m - m / 2) : 0.0f;
        all_ok = all_ok && ok[0] && ok[1] && ok[2];
        printf("%s{\"b\": %d, \"n\": %d, \"m\": %d, \"pruned_equals_oracle\": %s, \"cluster_equals_oracle\": %s, \"pruned_resumed_equals_oracle\": %s, "
               "\"rc\": [%d, %d, %d], \"pruned_ms\": %.4f, \"cluster_ms\": %.4f, \"pruned_ns_per_round\": %.1f, \"pruned_best_ms\": %.4f, \"pruned_half_best_ms\": %.4f, "
               "\"pruned_late_round_ns\": %.1f, \"last_error\": \"%s\"}",
               ci ? ", " : "", b, n, m, ok[0] ? "true" : "false", ok[1] ? "true" : "false", ok[2] ? "true" : "false", rc[0], rc[1], rc[2],
               ms_bucket, ms_cluster, 1e6 * ms_bucket / (m > 1 ? m - 1 : 1), best_full, best_half, late_round_ns, (rc[0] | rc[1] | rc[2]) ? ssd3d_last_error() : "");
        cudaFree(d_pts); cudaFree(d_temp); cudaFree(d_out);
    }
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("], \"all_ok\": %s, \"gpu\": \"%s\"}\n", all_ok ? "true" : "false", prop.name);
    return all_ok ? 0 : 1;
}
