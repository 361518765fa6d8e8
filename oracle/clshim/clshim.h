/*
 * clshim.h -- TEST INFRASTRUCTURE ONLY.  A minimal OpenCL-C-on-CPU environment
 * that lets gcc compile the reference's kernel sources (kernels/nbody/*.cl)
 * where they lie under /root/reference and run them as ONE work-group of 16
 * work-items (the reference's WORKGROUP_SIZE, GPUBarnesHutNBodySimulation.java:47-49;
 * every kernel is a grid-stride loop, so the number of groups does not change
 * results).
 *
 * Work-items are ucontext fibres scheduled round-robin in the order 15,14,...,0
 * and switched only at barrier / mem_fence / work_group_all / atomic_load: that
 * reproduces the SIMD-lockstep behaviour the kernels rely on -- in
 * calculateforce.cl:125-131 every lane reads localPos[depth] before lane 0
 * increments it, which holds here because lane 0 always runs last in a round.
 * Nothing in the reference sources is modified or copied.
 */
#ifndef CLSHIM_H
#define CLSHIM_H
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>

typedef int atomic_int;
typedef float atomic_float;

void clshim_yield(void);
void clshim_barrier(void);
int clshim_all(int pred);
int clshim_lid(void);

#define __kernel
#define __global
#define __local static
#define local static
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2
enum { memory_order_relaxed, memory_order_acquire, memory_order_release, memory_order_acq_rel, memory_order_seq_cst };
enum { memory_scope_work_item, memory_scope_work_group, memory_scope_device };

#define get_local_id(d) (clshim_lid())
#define get_global_id(d) (clshim_lid())
#define get_group_id(d) (0)
#define get_num_groups(d) (1)
#define get_local_size(d) (WORKGROUP_SIZE)
#define get_global_size(d) (WORKGROUP_SIZE)
#define get_work_dim() (1)

#define barrier(flags) clshim_barrier()
#define mem_fence(flags) clshim_yield()
#define atomic_work_item_fence(flags, order, scope) ((void)0)
#define work_group_all(p) clshim_all((p) ? 1 : 0)
#define atomic_load_explicit(p, order, scope) (clshim_yield(), *(p))
#define atomic_store_explicit(p, v, order, scope) ((void)(*(p) = (v)))

static inline int atom_cmpxchg(volatile int *p, int cmp, int val) { int old = *p; if (old == cmp) *p = val; return old; }
static inline int atom_dec(volatile int *p) { int old = *p; *p = old - 1; return old; }
static inline int atom_inc(volatile int *p) { int old = *p; *p = old + 1; return old; }
static inline int atom_max(volatile int *p, int v) { int old = *p; if (v > old) *p = v; return old; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float rsqrt(float x) { return 1.0f / sqrtf(x); }
#undef fmin
#undef fmax
#define fmin(a, b) fminf((a), (b))
#define fmax(a, b) fmaxf((a), (b))

#define CLSHIM_ARGS                                                                                                         \
    float *posX, float *posY, float *posZ, float *velX, float *velY, float *velZ, float *accX, float *accY, float *accZ, \
        int *step, int *blockCount, int *bodyCount, float *radius, int *maxDepth, int *bottom, float *mass, int *child,  \
        int *start, int *sorted, int *error
#endif
