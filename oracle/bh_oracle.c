/*
 * bh_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * six-kernel Barnes-Hut step.  Nothing in the product path (gpu_nbody_b200/,
 * include/, libbhstep.so) links, imports or executes this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function restates the *semantics* of one OpenCL kernel of the reference
 * (paths relative to /root/reference):
 *
 *   bho_bounding_box   kernels/nbody/boundingbox.cl:21-199
 *   bho_build_tree     kernels/nbody/buildtree.cl:13-200
 *   bho_summarize      kernels/nbody/summarizetree.cl:16-179
 *   bho_sort           kernels/nbody/sort.cl:13-76
 *   bho_calculate_force kernels/nbody/calculateforce.cl:26-188
 *   bho_integrate      kernels/nbody/integrate.cl:18-70
 *   bho_step           src/ch/fhnw/woipv/nbody/simulation/gpu/GPUBarnesHutNBodySimulation.java:249-271
 *   bho_number_of_nodes  ...GPUBarnesHutNBodySimulation.java:219-227
 *
 * The reference kernels are written for concurrently scheduled work-items that
 * spin on each other (summarizetree.cl:122-150, sort.cl:36-39, buildtree.cl:93),
 * so a run-to-completion emulation of one work-item at a time would deadlock.
 * The restatement is therefore sequential where the result does not depend on
 * scheduling: sequential insertion (the octree shape is insertion-order
 * independent), ascending-index summarise (children are allocated after their
 * parent, i.e. at lower indices), descending-index sort, one vote group of
 * `vote_width` consecutive sorted bodies at a time for the force walk.
 *
 * Pinning: the restatement is checked against the reference's *own* kernel
 * sources executed under oracle/clshim (a work-item emulator that compiles the
 * .cl files where they lie); see oracle/clshim/README and tests/golden/.
 *
 * Floating-point policy.  The reference builds its kernels with MAD enabled
 * (GPUBarnesHutNBodySimulation.java:210), so whether `a*b+c` is fused is the
 * OpenCL compiler's choice.  `fma_policy` selects one of the two allowed
 * outcomes: 0 = never fused (separate IEEE mul and add), 1 = every source-level
 * `x*y + z` fused into one fmaf -- the policy the CUDA kernels use.  This file
 * must be compiled with -ffp-contract=off so that only the explicit fmaf calls
 * fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BHO_MAXDEPTH 64 /* calculateforce.cl:12 */

typedef struct bho_state {
    /* the reference's 20 buffers, in kernel-argument order (GPUBH:198-205) */
    float *posX, *posY, *posZ;   /* [m+1] bodies 0..n-1, cells n..m (root = m) */
    float *velX, *velY, *velZ;   /* [m+1] */
    float *accX, *accY, *accZ;   /* [m+1] */
    int32_t *step;               /* [1] init -1 */
    int32_t *blockCount;         /* [1] */
    int32_t *bodyCount;          /* [m+1] */
    float *radius;               /* [1] */
    int32_t *maxDepth;           /* [1] init 1, running max */
    int32_t *bottom;             /* [1] */
    float *mass;                 /* [m+1] */
    int32_t *child;              /* [8(m+1)] */
    int32_t *start;              /* [m+1] */
    int32_t *sorted;             /* [m+1] */
    int32_t *error;              /* [1] */
    /* compile-time constants of the reference turned into parameters */
    int32_t n;                   /* NBODIES */
    int32_t m;                   /* NUMBER_OF_NODES */
    float theta_macro;           /* the THETA macro = theta^2 (calculateforce.cl:15-16) */
    float epsilon;               /* EPSILON, added to r^2 (calculateforce.cl:14) */
    float timestep;              /* TIMESTEP (integrate.cl:12) */
    int32_t vote_width;          /* WARPSIZE / WORKGROUP_SIZE = 16 (GPUBH:47-49) */
    int32_t fma_policy;          /* see header */
    /* diagnostics written by bho_calculate_force */
    int64_t interactions;        /* sum over bodies of (body,node) force evaluations */
    int64_t opens;               /* sum over bodies of (body,cell) opening tests that pushed */
    int32_t *group_interactions; /* optional [ceil(n/vote_width)]: node evaluations per vote group */
} bho_state;

/* GPUBarnesHutNBodySimulation.java:219-227 with maxComputeUnits = 16 (:126), WARPSIZE = 16 (:47) */
int32_t bho_number_of_nodes(int32_t nbodies) {
    int32_t nodes = nbodies * 2;
    if (nodes < 1024 * 16) nodes = 1024 * 16;
    while ((nodes & 15) != 0) ++nodes;
    return nodes;
}

/* boundingbox.cl:44-199.  fmin/fmax over all bodies seeded with body 0; the
 * result is independent of the reduction order. */
void bho_bounding_box(bho_state *s) {
    const int n = s->n, m = s->m;
    float minX = s->posX[0], minY = s->posY[0], minZ = s->posZ[0];
    float maxX = minX, maxY = minY, maxZ = minZ;
    for (int i = 0; i < n; ++i) {
        minX = fminf(minX, s->posX[i]); maxX = fmaxf(maxX, s->posX[i]);
        minY = fminf(minY, s->posY[i]); maxY = fmaxf(maxY, s->posY[i]);
        minZ = fminf(minZ, s->posZ[i]); maxZ = fmaxf(maxZ, s->posZ[i]);
    }
    const float rootX = 0.5f * (minX + maxX);          /* :171-173 */
    const float rootY = 0.5f * (minY + maxY);
    const float rootZ = 0.5f * (minZ + maxZ);
    *s->radius = 0.5f * fmaxf(fmaxf(maxX - minX, maxY - minY), maxZ - minZ); /* :179 */
    *s->bottom = m;                                    /* :180 */
    *s->blockCount = 0;                                /* :181 */
    s->posX[m] = rootX; s->posY[m] = rootY; s->posZ[m] = rootZ; /* :185-187 */
    s->mass[m] = -1.0f;                                /* :188 */
    s->start[m] = 0;                                   /* :189 */
    for (int i = 0; i < 8; ++i) s->child[8 * (int64_t)m + i] = -1; /* :193 */
    (*s->step)++;                                      /* :195 */
}

/* buildtree.cl:42-199, one body at a time.  Returns 0 or 1 (= *error). */
int32_t bho_build_tree(bho_state *s) {
    const int n = s->n, m = s->m;
    int32_t *child = s->child;
    const float radius = *s->radius;
    const float rootX = s->posX[m], rootY = s->posY[m], rootZ = s->posZ[m];
    int localMaxDepth = 1;
    for (int body = 0; body < n; ++body) {
        const float bx = s->posX[body], by = s->posY[body], bz = s->posZ[body];
        int node = m, depth = 1;
        float r = radius;
        int path = 0;                                   /* :65-68, strict < */
        if (rootX < bx) path = 1;
        if (rootY < by) path += 2;
        if (rootZ < bz) path += 4;
        int ch = child[8 * (int64_t)node + path];
        while (ch >= n) {                               /* :77-89 */
            node = ch; ++depth; r *= 0.5f;
            path = 0;
            if (s->posX[node] < bx) path = 1;
            if (s->posY[node] < by) path += 2;
            if (s->posZ[node] < bz) path += 4;
            ch = child[8 * (int64_t)node + path];
        }
        if (ch == -1) {                                 /* :98-101 */
            child[8 * (int64_t)node + path] = body;
        } else {                                        /* :102-180 */
            const int64_t locked = 8 * (int64_t)node + path;
            int patch = -1;
            do {
                depth++;
                const int cell = (*s->bottom)-- - 1;     /* atom_dec(_bottom) - 1, :109 */
                if (cell <= n) {                         /* :112-119 */
                    *s->error = 1;
                    *s->bottom = m;
                    return 1;
                }
                if (cell > patch) patch = cell;
                float x = (float)(path & 1) * r;         /* :124-126 */
                float y = (float)((path >> 1) & 1) * r;
                float z = (float)((path >> 2) & 1) * r;
                r *= 0.5f;
                s->mass[cell] = -1.0f;                   /* :131-132 */
                s->start[cell] = -1;
                x = s->posX[cell] = s->posX[node] - r + x; /* :134-136, left to right */
                y = s->posY[cell] = s->posY[node] - r + y;
                z = s->posZ[cell] = s->posZ[node] - r + z;
                for (int k = 0; k < 8; ++k) child[8 * (int64_t)cell + k] = -1;
                if (patch != cell) child[8 * (int64_t)node + path] = cell; /* :141-146 */
                path = 0;                                /* :148-152: re-insert the old body */
                if (x < s->posX[ch]) path = 1;
                if (y < s->posY[ch]) path += 2;
                if (z < s->posZ[ch]) path += 4;
                child[8 * (int64_t)cell + path] = ch;
                node = cell;                             /* :155-161: octant of the new body */
                path = 0;
                if (x < bx) path = 1;
                if (y < by) path += 2;
                if (z < bz) path += 4;
                ch = child[8 * (int64_t)node + path];
            } while (ch >= 0);
            child[8 * (int64_t)node + path] = body;      /* :169 */
            child[locked] = patch;                       /* :180 */
        }
        if (depth > localMaxDepth) localMaxDepth = depth; /* :186 */
    }
    if (localMaxDepth > *s->maxDepth) *s->maxDepth = localMaxDepth; /* atom_max, :199 */
    return 0;
}

/* buildtree.cl:42-199 as the reference runs it: bodies inserted CONCURRENTLY (OpenMP threads in place of work-items),
 * a child slot locked with compare-and-swap to -2 while a leaf is split (:93-101), cells taken from `bottom` with an atomic
 * decrement (:109), the finished sub-tree published into the locked slot after a fence (:173-180).  The tree SHAPE is the
 * one bho_build_tree produces; cell numbers depend on the race, as in the reference.  Used for the CPU baseline
 * (bench.py), where a single-threaded build would be a third of the step; the tests use both. */
int32_t bho_build_tree_parallel(bho_state *s) {
    const int n = s->n, m = s->m;
    int32_t *child = s->child;
    const float radius = *s->radius;
    const float rootX = s->posX[m], rootY = s->posY[m], rootZ = s->posZ[m];
    int globalMaxDepth = 1, failed = 0;
#pragma omp parallel for schedule(dynamic, 2048) reduction(max : globalMaxDepth)
    for (int body = 0; body < n; ++body) {
        if (__atomic_load_n(&failed, __ATOMIC_RELAXED)) continue;
        const float bx = s->posX[body], by = s->posY[body], bz = s->posZ[body];
        int node = m, depth = 1;
        float r = radius;
        int path = (rootX < bx ? 1 : 0) + (rootY < by ? 2 : 0) + (rootZ < bz ? 4 : 0);   /* :65-68, strict < */
        for (;;) {
            int32_t *slot = &child[8 * (int64_t)node + path];
            int ch = __atomic_load_n(slot, __ATOMIC_ACQUIRE);
            while (ch >= n) {                               /* :77-89 */
                node = ch; ++depth; r *= 0.5f;
                path = (s->posX[node] < bx ? 1 : 0) + (s->posY[node] < by ? 2 : 0) + (s->posZ[node] < bz ? 4 : 0);
                slot = &child[8 * (int64_t)node + path];
                ch = __atomic_load_n(slot, __ATOMIC_ACQUIRE);
            }
            if (ch == -2) {                                 /* locked by another inserter: look again (:93) */
                if (__atomic_load_n(&failed, __ATOMIC_RELAXED)) break;
                continue;
            }
            int expected = ch;
            if (!__atomic_compare_exchange_n(slot, &expected, -2, 0, __ATOMIC_ACQUIRE, __ATOMIC_RELAXED)) continue;
            if (ch == -1) {                                 /* :98-101 */
                __atomic_store_n(slot, body, __ATOMIC_RELEASE);
                break;
            }
            int patch = -1, ok = 1;                         /* :102-180 */
            int cur = node, curPath = path;
            do {
                depth++;
                const int cell = __atomic_fetch_sub(s->bottom, 1, __ATOMIC_RELAXED) - 1;   /* :109 */
                if (cell <= n) {                            /* :112-119 */
                    __atomic_store_n(&failed, 1, __ATOMIC_RELAXED);
                    ok = 0;
                    break;
                }
                if (cell > patch) patch = cell;
                float x = (float)(curPath & 1) * r;          /* :124-126 */
                float y = (float)((curPath >> 1) & 1) * r;
                float z = (float)((curPath >> 2) & 1) * r;
                r *= 0.5f;
                s->mass[cell] = -1.0f;                      /* :131-132 */
                s->start[cell] = -1;
                x = s->posX[cell] = s->posX[cur] - r + x;    /* :134-136, left to right */
                y = s->posY[cell] = s->posY[cur] - r + y;
                z = s->posZ[cell] = s->posZ[cur] - r + z;
                for (int k = 0; k < 8; ++k) child[8 * (int64_t)cell + k] = -1;
                if (patch != cell) child[8 * (int64_t)cur + curPath] = cell;   /* :141-146 */
                const int qp = (x < s->posX[ch] ? 1 : 0) + (y < s->posY[ch] ? 2 : 0) + (z < s->posZ[ch] ? 4 : 0);   /* :148-152 */
                child[8 * (int64_t)cell + qp] = ch;
                cur = cell;                                 /* :155-161 */
                curPath = (x < bx ? 1 : 0) + (y < by ? 2 : 0) + (z < bz ? 4 : 0);
            } while (child[8 * (int64_t)cur + curPath] >= 0);
            if (ok) {
                child[8 * (int64_t)cur + curPath] = body;   /* :169 */
                __atomic_store_n(slot, patch, __ATOMIC_RELEASE);   /* :173-180 */
            } else {
                __atomic_store_n(slot, ch, __ATOMIC_RELEASE);      /* give the leaf back; the error is reported below */
            }
            break;
        }
        if (depth > globalMaxDepth) globalMaxDepth = depth; /* :186 */
    }
    if (failed) {
        *s->error = 1;
        *s->bottom = m;
        return 1;
    }
    if (globalMaxDepth > *s->maxDepth) *s->maxDepth = globalMaxDepth; /* atom_max, :199 */
    return 0;
}

/* summarizetree.cl:55-178 in ascending cell order; children summed in octant
 * order (the reference's own order is timing dependent: ready children in
 * octant order, late ones in reverse arrival order, :93-111 vs :124-150). */
void bho_summarize(bho_state *s) {
    const int n = s->n, m = s->m, fma = s->fma_policy;
    int32_t *child = s->child;
    for (int node = *s->bottom; node <= m; ++node) {
        float cellMass = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
        int count = 0, used = 0;
        for (int k = 0; k < 8; ++k) {
            const int ch = child[8 * (int64_t)node + k];
            if (ch < 0) continue;
            if (k != used) {                             /* :77-81 compaction */
                child[8 * (int64_t)node + k] = -1;
                child[8 * (int64_t)node + used] = ch;
            }
            const float mch = s->mass[ch];
            if (ch >= n) count += s->bodyCount[ch] - 1;  /* :98-105 */
            cellMass += mch;                             /* :107-110 */
            if (fma) {
                cx = fmaf(s->posX[ch], mch, cx);
                cy = fmaf(s->posY[ch], mch, cy);
                cz = fmaf(s->posZ[ch], mch, cz);
            } else {
                cx += s->posX[ch] * mch;
                cy += s->posY[ch] * mch;
                cz += s->posZ[ch] * mch;
            }
            used++;
        }
        count += used;                                   /* :118 */
        s->bodyCount[node] = count;                      /* :160 */
        const float inv = 1.0f / cellMass;               /* :161 */
        s->posX[node] = cx * inv;                        /* :165-167 */
        s->posY[node] = cy * inv;
        s->posZ[node] = cz * inv;
        s->mass[node] = cellMass;                        /* :172 */
    }
}

/* sort.cl:26-72 in descending cell order (parents before children). */
void bho_sort(bho_state *s) {
    const int n = s->n, m = s->m;
    for (int cell = m; cell >= *s->bottom; --cell) {
        int st = s->start[cell];
        for (int i = 0; i < 8; ++i) {
            const int ch = s->child[8 * (int64_t)cell + i];
            if (ch >= n) {                               /* :44-56 */
                s->start[ch] = st;
                st += s->bodyCount[ch];
            } else if (ch >= 0) {                        /* :59-65 */
                s->sorted[st] = ch;
                ++st;
            }
        }
    }
}

/* calculateforce.cl:52-67.  dq[d] = radius^2 * 0.25^d / THETA + EPSILON. */
static int bho_fill_dq(const bho_state *s, float *dq) {
    const float radius = *s->radius;
    const int maxDepth = *s->maxDepth;
    if (maxDepth > BHO_MAXDEPTH) return 1;               /* :69-73 */
    dq[0] = (s->theta_macro > 0) ? radius * radius / s->theta_macro : radius * radius;
    int i;
    for (i = 1; i < maxDepth; ++i) {
        dq[i] = 0.25f * dq[i - 1];
        dq[i - 1] += s->epsilon;
    }
    dq[i - 1] += s->epsilon;
    return 0;
}

/* calculateforce.cl:99-185 for one vote group [k0, k0+nl). */
static void bho_force_group(bho_state *s, const float *dq, int k0, int nl, int fma,
                            int64_t *inter_out, int64_t *open_out) {
    const int n = s->n, m = s->m;
    const float eps = s->epsilon, dt = s->timestep;
    const int32_t *child = s->child;
    float px[64], py[64], pz[64], ax[64], ay[64], az[64], dx[64], dy[64], dz[64], r2[64];
    int idx[64];
    for (int l = 0; l < nl; ++l) {
        idx[l] = s->sorted[k0 + l];                      /* :101 */
        px[l] = s->posX[idx[l]]; py[l] = s->posY[idx[l]]; pz[l] = s->posZ[idx[l]];
        ax[l] = ay[l] = az[l] = 0.0f;
    }
    int stackNode[BHO_MAXDEPTH + 1], stackPos[BHO_MAXDEPTH + 1];
    int depth = 0;
    int64_t inter = 0, opens = 0;
    stackNode[0] = m; stackPos[0] = 0;                   /* :113-117 */
    while (depth >= 0) {                                 /* :122 */
        int top;
        while ((top = stackPos[depth]) < 8) {            /* :125 */
            const int ch = child[8 * (int64_t)stackNode[depth] + top];
            stackPos[depth] = top + 1;                   /* :128-131 */
            if (ch >= 0) {
                const float cx = s->posX[ch], cy = s->posY[ch], cz = s->posZ[ch];
                int all = 1;
                const float thr = dq[depth];
                for (int l = 0; l < nl; ++l) {           /* :138-143 */
                    dx[l] = cx - px[l]; dy[l] = cy - py[l]; dz[l] = cz - pz[l];
                    if (fma) r2[l] = fmaf(dz[l], dz[l], fmaf(dy[l], dy[l], dx[l] * dx[l])) + eps;
                    else r2[l] = dx[l] * dx[l] + dy[l] * dy[l] + dz[l] * dz[l] + eps;
                    all &= (r2[l] >= thr);
                }
                if (ch < n || all) {                     /* :145 */
                    const float mc = s->mass[ch];
                    for (int l = 0; l < nl; ++l) {       /* :146-151 */
                        const float rinv = 1.0f / sqrtf(r2[l]);
                        const float f = mc * rinv * rinv * rinv;
                        if (fma) {
                            ax[l] = fmaf(dx[l], f, ax[l]);
                            ay[l] = fmaf(dy[l], f, ay[l]);
                            az[l] = fmaf(dz[l], f, az[l]);
                        } else {
                            ax[l] += dx[l] * f; ay[l] += dy[l] * f; az[l] += dz[l] * f;
                        }
                    }
                    inter += nl;
                } else {                                 /* :154-163 push */
                    depth++;
                    stackNode[depth] = ch; stackPos[depth] = 0;
                    opens += nl;
                }
            } else {
                depth = depth - 1 > 0 ? depth - 1 : 0;   /* :166 */
            }
        }
        depth--;                                         /* :171 */
    }
    for (int l = 0; l < nl; ++l) {
        const int b = idx[l];
        if (*s->step > 0) {                              /* :174-179 */
            s->velX[b] += (ax[l] - s->accX[b]) * dt * 0.5f;
            s->velY[b] += (ay[l] - s->accY[b]) * dt * 0.5f;
            s->velZ[b] += (az[l] - s->accZ[b]) * dt * 0.5f;
        }
        s->accX[b] = ax[l]; s->accY[b] = ay[l]; s->accZ[b] = az[l]; /* :183-185 */
    }
    *inter_out = inter; *open_out = opens;
}

/* Force walk for the sorted slots [first, first+count); first must be a
 * multiple of vote_width.  (0, n) is the whole kernel. */
int32_t bho_calculate_force_range(bho_state *s, int32_t first, int32_t count) {
    float dq[BHO_MAXDEPTH + 1];
    if (bho_fill_dq(s, dq)) { *s->error = 1; return 1; }
    const int w = s->vote_width;
    const int g0 = first / w, g1 = (first + count + w - 1) / w;
    int64_t tin = 0, top = 0;
    const int fma = s->fma_policy;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : tin, top)
    for (int g = g0; g < g1; ++g) {
        const int k0 = g * w;
        int nl = first + count - k0;
        if (nl > w) nl = w;
        int64_t a = 0, b = 0;
        bho_force_group(s, dq, k0, nl, fma, &a, &b);
        if (s->group_interactions) s->group_interactions[g] = (int32_t)(a / nl);
        tin += a; top += b;
    }
    s->interactions = tin; s->opens = top;
    return 0;
}

int32_t bho_calculate_force(bho_state *s) { return bho_calculate_force_range(s, 0, s->n); }

/* integrate.cl:27-43 */
void bho_integrate(bho_state *s) {
    const int n = s->n, fma = s->fma_policy;
    const float dt = s->timestep;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const float dvx = s->accX[i] * dt * 0.5f, dvy = s->accY[i] * dt * 0.5f, dvz = s->accZ[i] * dt * 0.5f;
        const float vx = s->velX[i] + dvx, vy = s->velY[i] + dvy, vz = s->velZ[i] + dvz;
        if (fma) {
            s->posX[i] = fmaf(vx, dt, s->posX[i]);
            s->posY[i] = fmaf(vy, dt, s->posY[i]);
            s->posZ[i] = fmaf(vz, dt, s->posZ[i]);
        } else {
            s->posX[i] += vx * dt; s->posY[i] += vy * dt; s->posZ[i] += vz * dt;
        }
        s->velX[i] = vx + dvx; s->velY[i] = vy + dvy; s->velZ[i] = vz + dvz;
    }
}

/* GPUBarnesHutNBodySimulation.java:258-263 */
int32_t bho_step(bho_state *s, int32_t nsteps) {
    for (int i = 0; i < nsteps; ++i) {
        bho_bounding_box(s);
        if (bho_build_tree(s)) return 1;
        bho_summarize(s);
        bho_sort(s);
        if (bho_calculate_force(s)) return 1;
        bho_integrate(s);
    }
    return 0;
}

/* ------------------------------------------------------------------------- *
 * Comparators (not part of the reference; test helpers)
 * ------------------------------------------------------------------------- */

/* Canonical relabelling: the reference's cell numbers depend on the atom_dec
 * race (buildtree.cl:109), the tree shape does not.  Walk the tree depth-first
 * from the root following child slots 0..7 and number cells in visit order.
 * order[c]      = original index of the c-th visited cell (order[0] = m)
 * canon[8c+k]   = -1 | body index | n + (visit number of the child cell)
 * Returns the number of cells visited, or -1 if more than max_cells or a cycle
 * / out-of-range index is found. */
int32_t bho_canonicalize(const int32_t *child, int32_t n, int32_t m, int32_t max_cells,
                         int32_t *order, int32_t *canon) {
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (8 * (size_t)BHO_MAXDEPTH * 4 + 16));
    const int cap = 8 * BHO_MAXDEPTH * 4 + 16;
    int sp = 0, visited = 0;
    /* explicit pre-order: a cell gets its number when popped; children pushed
     * in reverse slot order so that slot 0 is visited first. */
    stack[sp++] = m;
    /* visit numbers must be known when the parent row is written, so assign
     * numbers in a first pass (pre-order) and fill rows in a second. */
    int32_t *number = (int32_t *)malloc(sizeof(int32_t) * ((size_t)m - n + 1));
    memset(number, 0xff, sizeof(int32_t) * ((size_t)m - n + 1));
    while (sp > 0) {
        const int cell = stack[--sp];
        if (cell < n || cell > m || number[cell - n] != -1 || visited >= max_cells) {
            free(stack); free(number); return -1;
        }
        number[cell - n] = visited;
        order[visited++] = cell;
        for (int k = 7; k >= 0; --k) {
            const int ch = child[8 * (int64_t)cell + k];
            if (ch >= n) {
                if (sp >= cap) { free(stack); free(number); return -1; }
                stack[sp++] = ch;
            }
        }
    }
    for (int c = 0; c < visited; ++c) {
        const int cell = order[c];
        for (int k = 0; k < 8; ++k) {
            const int ch = child[8 * (int64_t)cell + k];
            canon[8 * (int64_t)c + k] = (ch >= n) ? n + number[ch - n] : ch;
        }
    }
    free(stack); free(number);
    return visited;
}

/* Total energy, double precision, same formula on both sides of every parity
 * test: sum 1/2 m v^2 - sum_{i<j} m_i m_j / sqrt(r^2 + eps).  (The reference's
 * own printEnergy, GPUBH:305-340, uses an unsoftened doubled potential and is
 * deliberately not restated.) */
void bho_energy(int32_t n, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                const float *vz, const float *mass, float eps, double *ekin_out, double *epot_out) {
    double ekin = 0.0, epot = 0.0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ekin, epot)
    for (int i = 0; i < n; ++i) {
        ekin += 0.5 * mass[i] * ((double)vx[i] * vx[i] + (double)vy[i] * vy[i] + (double)vz[i] * vz[i]);
        double p = 0.0;
        for (int j = i + 1; j < n; ++j) {
            const double dx = (double)x[j] - x[i], dy = (double)y[j] - y[i], dz = (double)z[j] - z[i];
            p += mass[j] / sqrt(dx * dx + dy * dy + dz * dz + eps);
        }
        epot -= mass[i] * p;
    }
    *ekin_out = ekin; *epot_out = epot;
}

/* Direct O(n*count) softened sum for bodies [first, first+count) in double:
 * physics sanity check for the tree code (not a reference kernel). */
void bho_direct_acc(int32_t n, const float *x, const float *y, const float *z, const float *mass, float eps,
                    int32_t first, int32_t count, double *ax, double *ay, double *az) {
#pragma omp parallel for schedule(static)
    for (int i = first; i < first + count; ++i) {
        double sx = 0, sy = 0, sz = 0;
        for (int j = 0; j < n; ++j) {
            const double dx = (double)x[j] - x[i], dy = (double)y[j] - y[i], dz = (double)z[j] - z[i];
            const double r2 = dx * dx + dy * dy + dz * dz + eps;
            const double f = mass[j] / (r2 * sqrt(r2));
            sx += dx * f; sy += dy * f; sz += dz * f;
        }
        ax[i - first] = sx; ay[i - first] = sy; az[i - first] = sz;
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline legs ask for all host cores explicitly. */
void bho_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int32_t bho_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
