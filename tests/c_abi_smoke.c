/* Plain-C user of include/bhstep.h: proves that the boundary is a C ABI (compiled with gcc -std=c99, no C++),
 * loads libbhstep.so with dlopen like a JNI/FFM/cgo binding would, and runs what can run without a GPU.
 * With a GPU (argv[1] = "gpu") it also runs two steps of a small universe and checks a few invariants. */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bhstep.h"

#define SYM(name) name##_t p_##name = (name##_t)dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }
typedef int (*bh_create_t)(bh_sim **, int32_t, float, float, float, int32_t, int32_t);
typedef void (*bh_destroy_t)(bh_sim *);
typedef const char *(*bh_last_error_t)(bh_sim *);
typedef int32_t (*bh_number_of_nodes_t)(int32_t);
typedef int32_t (*bh_abi_version_t)(void);
typedef int (*bh_upload_t)(bh_sim *, const float *, const float *, const float *, const float *, const float *, const float *, const float *);
typedef int (*bh_step_t)(bh_sim *, int32_t);
typedef int (*bh_read_t)(bh_sim *, int32_t, void *, int64_t);
typedef int (*bh_stats_t_fn)(bh_sim *, bh_stats_t *);

int main(int argc, char **argv) {
    void *lib = dlopen(argc > 2 ? argv[2] : "libbhstep.so", RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(bh_create) SYM(bh_destroy) SYM(bh_last_error) SYM(bh_number_of_nodes) SYM(bh_abi_version) SYM(bh_upload) SYM(bh_step) SYM(bh_read)
    bh_stats_t_fn p_bh_stats = (bh_stats_t_fn)dlsym(lib, "bh_stats");
    if (!p_bh_stats) return 2;
    if (p_bh_abi_version() != 2) return 3;
    if (p_bh_number_of_nodes(32768) != 65536 || p_bh_number_of_nodes(1) != 16384) return 4;  /* GPUBH:219-227 */
    const int gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    bh_sim *sim = NULL;
    const int n = 4096;
    int rc = p_bh_create(&sim, n, 0.5f, 0.0025f, 0.025f, 16, 0);
    if (!gpu) {
        if (rc != BH_ERR_NO_DEVICE && rc != BH_OK) { fprintf(stderr, "unexpected rc %d\n", rc); return 5; }
        if (rc == BH_ERR_NO_DEVICE && !strstr(p_bh_last_error(NULL), "no CPU fallback")) return 6;
        if (sim) p_bh_destroy(sim);
        printf("c_abi_smoke ok (rc=%d)\n", rc);
        return 0;
    }
    if (rc != BH_OK) { fprintf(stderr, "bh_create: %d %s\n", rc, p_bh_last_error(NULL)); return 7; }
    float *a[7];
    for (int k = 0; k < 7; ++k) a[k] = (float *)calloc(n, sizeof(float));
    unsigned s = 12345u;
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) { s = s * 1664525u + 1013904223u; a[k][i] = ((s >> 8) / 16777216.0f - 0.5f) * 6.0f; }
        a[6][i] = 1.0f / n;
    }
    if ((rc = p_bh_upload(sim, a[0], a[1], a[2], a[3], a[4], a[5], a[6])) != 0) return 8;
    if ((rc = p_bh_step(sim, 2)) != 0) { fprintf(stderr, "bh_step: %d %s\n", rc, p_bh_last_error(sim)); return 9; }
    bh_stats_t st;
    if (p_bh_stats(sim, &st) != 0 || st.step != 1 || st.error != 0 || st.cells_used < n / 8 || st.cells_used > n) return 10;
    int32_t *sorted = (int32_t *)malloc(sizeof(int32_t) * n);
    char *seen = (char *)calloc(n, 1);
    if (p_bh_read(sim, BH_SORTED, sorted, n) != 0) return 11;
    for (int i = 0; i < n; ++i) { if (sorted[i] < 0 || sorted[i] >= n || seen[sorted[i]]) return 12; seen[sorted[i]] = 1; }
    p_bh_destroy(sim);
    printf("c_abi_smoke ok (gpu: %d cells, depth %d)\n", st.cells_used, st.max_depth);
    return 0;
}
