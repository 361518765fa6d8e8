/*
 * ref_driver.c -- TEST INFRASTRUCTURE ONLY.  Runs the reference's own OpenCL
 * kernel sources (compiled by gcc through clshim.h) the way
 * GPUBarnesHutNBodySimulation.java does: buffers as in loadBuffers (:153-181),
 * the six kernels in step() order (:258-263), one work-group of WORKGROUP_SIZE
 * work-items.  Dumps all 20 buffers after a requested kernel.
 *
 *   ref_step <universe.bin> <out.bin> <steps> [stop_after_kernel 0..5 of the last step]
 *
 * universe.bin: 7 x NBODIES float32 (x y z vx vy vz mass), native endianness.
 * out.bin: the 20 buffers in kernel-argument order, each with its full length.
 * NBODIES, NUMBER_OF_NODES, WORKGROUP_SIZE, NUM_WORK_GROUPS are compile-time
 * macros, as in the reference's build options (:207-217).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#ifndef WORKGROUP_SIZE
#error "compile with -DWORKGROUP_SIZE=16 -DNUM_WORK_GROUPS=1 -DNBODIES=.. -DNUMBER_OF_NODES=.."
#endif

#define NWI WORKGROUP_SIZE
#define STACK_BYTES (1 << 18)

typedef void (*kernel_fn)();
void boundingBox();
void buildTree();
void summarizeTree();
void sort();
void calculateForce();
void integrate();

/* ---- fibre scheduler ------------------------------------------------------ */
static ucontext_t sched_ctx, wi_ctx[NWI];
static char *wi_stack[NWI];
static int wi_done[NWI];
static int cur = -1, live = 0;
/* barrier / vote state */
static int bar_arrived = 0, bar_generation = 0;
static int vote_arrived = 0, vote_and = 1, vote_result = 0, vote_generation = 0;

static void *kargs[20];
static kernel_fn kcur;

int clshim_lid(void) { return cur; }

void clshim_yield(void) { swapcontext(&wi_ctx[cur], &sched_ctx); }

void clshim_barrier(void) {
    const int gen = bar_generation;
    bar_arrived++;
    while (gen == bar_generation) {
        if (bar_arrived >= live) {  /* everybody still running has arrived */
            bar_arrived = 0;
            bar_generation++;
        }
        clshim_yield(); /* also after completing: everybody restarts in lane order 15..0 */
    }
}

int clshim_all(int pred) {
    const int gen = vote_generation;
    vote_and &= pred;
    vote_arrived++;
    while (gen == vote_generation) {
        if (vote_arrived >= live) {
            vote_result = vote_and;
            vote_arrived = 0;
            vote_and = 1;
            vote_generation++;
        }
        clshim_yield(); /* also after completing: lane 0 must not run ahead of the others */
    }
    return vote_result;
}

static void wi_entry(void) {
    kcur(kargs[0], kargs[1], kargs[2], kargs[3], kargs[4], kargs[5], kargs[6], kargs[7], kargs[8], kargs[9], kargs[10], kargs[11],
         kargs[12], kargs[13], kargs[14], kargs[15], kargs[16], kargs[17], kargs[18], kargs[19]);
    wi_done[cur] = 1;
    live--;
    /* a work-item that leaves while others wait at a barrier / vote must not block them */
    swapcontext(&wi_ctx[cur], &sched_ctx);
}

static void run_kernel(kernel_fn k) {
    kcur = k;
    live = NWI;
    bar_arrived = vote_arrived = 0;
    vote_and = 1;
    for (int i = 0; i < NWI; ++i) {
        wi_done[i] = 0;
        getcontext(&wi_ctx[i]);
        wi_ctx[i].uc_stack.ss_sp = wi_stack[i];
        wi_ctx[i].uc_stack.ss_size = STACK_BYTES;
        wi_ctx[i].uc_link = &sched_ctx;
        makecontext(&wi_ctx[i], wi_entry, 0);
    }
    while (live > 0) {
        /* highest lane first, lane 0 last: see clshim.h */
        for (int i = NWI - 1; i >= 0; --i) {
            if (wi_done[i]) continue;
            cur = i;
            swapcontext(&sched_ctx, &wi_ctx[i]);
        }
        /* a barrier/vote whose last participant exited instead of arriving */
        if (live > 0 && bar_arrived >= live && bar_arrived > 0) { bar_arrived = 0; bar_generation++; }
        if (live > 0 && vote_arrived >= live && vote_arrived > 0) { vote_result = vote_and; vote_arrived = 0; vote_and = 1; vote_generation++; }
    }
    cur = -1;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s universe.bin out.bin steps [stop_after_kernel]\n", argv[0]);
        return 2;
    }
    const int steps = atoi(argv[3]);
    const int stop_after = argc > 4 ? atoi(argv[4]) : 5;
    const size_t n = NBODIES, m1 = (size_t)NUMBER_OF_NODES + 1;
    /* GPUBH:155-179 */
    float *f[10];
    for (int i = 0; i < 10; ++i) f[i] = calloc(m1, sizeof(float)); /* posXYZ velXYZ accXYZ mass */
    int *bodyCount = calloc(m1, sizeof(int)), *child = calloc(8 * m1, sizeof(int)), *start = calloc(m1, sizeof(int)),
        *sorted = calloc(m1, sizeof(int));
    int step = -1, blockCount = 0, maxDepth = 1, bottom = 0, error = 0;
    float radius = 0.0f;
    FILE *in = fopen(argv[1], "rb");
    if (!in) { perror(argv[1]); return 1; }
    const int order[7] = {0, 1, 2, 3, 4, 5, 9};
    for (int i = 0; i < 7; ++i)
        if (fread(f[order[i]], sizeof(float), n, in) != n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(in);
    void *args[20] = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], &step, &blockCount, bodyCount, &radius, &maxDepth,
                      &bottom, f[9], child, start, sorted, &error};
    memcpy(kargs, args, sizeof args);
    for (int i = 0; i < NWI; ++i) wi_stack[i] = malloc(STACK_BYTES);
    kernel_fn seq[6] = {boundingBox, buildTree, summarizeTree, sort, calculateForce, integrate}; /* GPUBH:258-263 */
    for (int s = 0; s < steps && !error; ++s)
        for (int k = 0; k < 6; ++k) {
            run_kernel(seq[k]);
            if (error) break;
            if (s == steps - 1 && k == stop_after) goto done;
        }
done:;
    FILE *out = fopen(argv[2], "wb");
    if (!out) { perror(argv[2]); return 1; }
    const size_t len[20] = {m1, m1, m1, m1, m1, m1, m1, m1, m1, 1, 1, m1, 1, 1, 1, m1, 8 * m1, m1, m1, 1};
    for (int i = 0; i < 20; ++i) fwrite(args[i], 4, len[i], out);
    fclose(out);
    return error ? 3 : 0;
}
