// bh_kernels.cuh -- hand-written sm_100a kernels of the Barnes-Hut step.
//
// One kernel per stage of the reference's step (GPUBarnesHutNBodySimulation.java:258-263):
//   bbox_kernel       <- kernels/nbody/boundingbox.cl
//   build_kernel      <- kernels/nbody/buildtree.cl
//   summarize_kernel  <- kernels/nbody/summarizetree.cl
//   sort_kernel       <- kernels/nbody/sort.cl
//   walk_kernel       <- kernels/nbody/calculateforce.cl   (deep_walk_kernel: fallback for very deep trees / 32-wide votes)
//   finish_kernel     <- calculateforce.cl:174-185 (velocity correction) + kernels/nbody/integrate.cl
// They reproduce the reference's *results* (see DESIGN.md for the parity classes),
// not its code: the data layout, work decomposition and synchronisation are
// designed for B200.
//
// HBM layout (N bodies, M = number of nodes, NC = M - N + 1 cell slots).  Bodies are
// stored PHYSICALLY in the previous step's tree (DFS) order: "slot" k holds the body
// that was k-th in sorted[] when the last step finished; the host's numbering of
// that body travels with it (origId).  All body arrays are double buffered; the
// fused finish kernel reads slot perm[k] of the current buffers and writes slot k
// of the other ones.
//   body4  float4[2][N]   {x, y, z, mass} per slot
//   velacc float4[2][2N]  {vx,vy,vz,origId (int bits)},{ax,ay,az,0} per slot: one 32-byte sector
//   cell4  float4[NC]     cell c at [c-N]; geometric centre and mass = -1 during build,
//                         {COM, mass} after summarise; root = cell M
//   child  int[8*NC]      child[(cell-N)*8 + k]; -1 empty, -2 locked, <N body slot, >=N cell
//   octet  float4[8*NC]   the force walk's record of a cell, written by summarise: copies of its
//   ometa  int2[8*NC]     children's {x,y,z,mass}, child cells first, then child bodies, and per child
//                         {opening threshold dq[level] as float bits, walk entry}: walk entry of a cell =
//                         (cell-N) | (#children-1) << 27; bodies have threshold -1 (always accepted) and entry -1
//   meta   int[NC]        #child cells | #child bodies << 4 (deep walk kernel)
//   start  int[NC]        `start` of the reference
//   count  int[NC]        bodyCount | (#children-1) << 28; -1 = not summarised yet
//   parent, arrived int[NC]  parent pointer; level << 16 | #child cells << 8 | reports received (counter-driven summarise)
//   perm   int[N]         sort's output: perm[k] = slot of the k-th body in tree order (sorted[] = origId[perm[k]])
//   acc    float4[2][N+pad] accelerations in tree order (written by the walk; the multi-GPU all-gather buffer)
//
// Floating-point policy (DESIGN.md "FMA policy"): every source-level x*y+z of the
// reference is one fmaf, everything else a separately rounded IEEE operation,
// spelled with intrinsics so that nvcc cannot re-associate.  The oracle's
// fma_policy=1 is the same policy, which makes every stage except the rsqrt in
// the force kernel bit-reproducible on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

namespace bh {

constexpr int kMaxDepth = 64;       // MAXDEPTH, calculateforce.cl:12
constexpr int kLock = -2;           // LOCK, buildtree.cl:8
constexpr int kSpinBudget = 1 << 24;  // polls before a device-side wait gives up (error = 2)
constexpr int kEntryMask = 0x7ffffff;  // cell - N of a walk entry (27 bits: NC <= 2^27)
constexpr int kCountMask = 0xfffffff;  // bodyCount of a count word (28 bits)

constexpr int kNothingDirty = 0x7fffffff;
struct Scalars {
    int step;         // init -1 (GPUBH:165)
    int blockCount;   // last-block-done ticket of bbox_kernel
    float radius;
    int maxDepth;     // init 1, running max (buildtree.cl:199)
    int bottom;
    int error;
    int walkTicket;   // next chunk of eight vote groups the force walk hands to a warp (reset before every walk)
    int rootEntry;    // walk entry of the root cell (written by summarise)
    int walkSpills;   // times a group's cell stack spilled to global memory in the last walk (diagnostic)
    int lowWater;     // lowest cell a build has allocated since the last reset, kNothingDirty before the first bbox
    unsigned long long interactions;
    unsigned long long opens;
};

// ---- memory-model helpers -------------------------------------------------
// Cross-thread data inside a kernel is read with strong (L2) loads; publication is a release store / release atomic
// (MEMBAR.ALL.GPU, no L1 invalidate -- __threadfence() is fence.sc = MEMBAR.SC.GPU + CCTL.IVALL on sm_100).
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int *p, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ int atom_add_acquire(int *p, int v) {
    int old;
    asm volatile("atom.acquire.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ int octant(float cx, float cy, float cz, float bx, float by, float bz) {
    // buildtree.cl:65-68: strict <, ties go to the low octant
    return (cx < bx ? 1 : 0) + (cy < by ? 2 : 0) + (cz < bz ? 4 : 0);
}

// ---- 1. bounding box --------------------------------------------------------
// boundingbox.cl: min/max over bodies, root cell, per-step resets.  Warp-shuffle
// reduction, one shared-memory hop per block, last-block-done combine.
constexpr int kBboxThreads = 512;

__device__ __forceinline__ void warp_minmax(float &mnx, float &mny, float &mnz, float &mxx, float &mxy, float &mxz) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
}

// (boundingbox.cl:44-58 seeds every lane with body 0; min/max do not depend on the seed, slot 0 is used)
__global__ void __launch_bounds__(kBboxThreads) bbox_kernel(const float4 *__restrict__ body4, float4 *__restrict__ cell4,
                                                            int *__restrict__ child, int *__restrict__ start,
                                                            int *__restrict__ count, int *__restrict__ arrived,
                                                            float *__restrict__ partials, Scalars *__restrict__ sc, int n, int m) {
    __shared__ float red[6][kBboxThreads / 32];
    __shared__ bool isLast;
    const float4 seed = body4[0];
    float mnx = seed.x, mny = seed.y, mnz = seed.z, mxx = seed.x, mxy = seed.y, mxz = seed.z;
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
        const float4 p0 = body4[i], p1 = body4[i + stride], p2 = body4[i + 2 * stride], p3 = body4[i + 3 * stride];
        mnx = fminf(fminf(mnx, p0.x), fminf(fminf(p1.x, p2.x), p3.x)); mxx = fmaxf(fmaxf(mxx, p0.x), fmaxf(fmaxf(p1.x, p2.x), p3.x));
        mny = fminf(fminf(mny, p0.y), fminf(fminf(p1.y, p2.y), p3.y)); mxy = fmaxf(fmaxf(mxy, p0.y), fmaxf(fmaxf(p1.y, p2.y), p3.y));
        mnz = fminf(fminf(mnz, p0.z), fminf(fminf(p1.z, p2.z), p3.z)); mxz = fmaxf(fmaxf(mxz, p0.z), fmaxf(fmaxf(p1.z, p2.z), p3.z));
    }
    for (; i < n; i += stride) {
        const float4 p = body4[i];
        mnx = fminf(mnx, p.x); mxx = fmaxf(mxx, p.x);
        mny = fminf(mny, p.y); mxy = fmaxf(mxy, p.y);
        mnz = fminf(mnz, p.z); mxz = fmaxf(mxz, p.z);
    }
    warp_minmax(mnx, mny, mnz, mxx, mxy, mxz);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mnz;
        red[3][warp] = mxx; red[4][warp] = mxy; red[5][warp] = mxz;
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int nw = kBboxThreads / 32;
        mnx = red[0][lane % nw]; mny = red[1][lane % nw]; mnz = red[2][lane % nw];
        mxx = red[3][lane % nw]; mxy = red[4][lane % nw]; mxz = red[5][lane % nw];
        warp_minmax(mnx, mny, mnz, mxx, mxy, mxz);
        if (lane == 0) {
            float *out = partials + 6 * blockIdx.x;
            out[0] = mnx; out[1] = mny; out[2] = mnz; out[3] = mxx; out[4] = mxy; out[5] = mxz;
            __threadfence();
            // atomicInc wraps to 0 at gridDim-1: blockCount is back to 0 for the next step (boundingbox.cl:181)
            isLast = (atomicInc(reinterpret_cast<unsigned *>(&sc->blockCount), gridDim.x - 1) == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!isLast || warp != 0) return;
    __threadfence();
    mnx = seed.x; mny = seed.y; mnz = seed.z; mxx = seed.x; mxy = seed.y; mxz = seed.z;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
        const float *in = partials + 6 * b;
        mnx = fminf(mnx, __ldcg(in + 0)); mny = fminf(mny, __ldcg(in + 1)); mnz = fminf(mnz, __ldcg(in + 2));
        mxx = fmaxf(mxx, __ldcg(in + 3)); mxy = fmaxf(mxy, __ldcg(in + 4)); mxz = fmaxf(mxz, __ldcg(in + 5));
    }
    warp_minmax(mnx, mny, mnz, mxx, mxy, mxz);
    if (lane < 8) child[8 * (size_t)(m - n) + lane] = -1;  // boundingbox.cl:193
    if (lane == 0) {
        // boundingbox.cl:171-195
        const float rx = __fmul_rn(0.5f, __fadd_rn(mnx, mxx));
        const float ry = __fmul_rn(0.5f, __fadd_rn(mny, mxy));
        const float rz = __fmul_rn(0.5f, __fadd_rn(mnz, mxz));
        sc->radius = __fmul_rn(0.5f, fmaxf(fmaxf(__fsub_rn(mxx, mnx), __fsub_rn(mxy, mny)), __fsub_rn(mxz, mnz)));
        // what the next reset has to clear: the root row written here and the cells of every build so far
        sc->lowWater = sc->lowWater == kNothingDirty ? m : min(sc->lowWater, sc->bottom);
        sc->bottom = m;
        cell4[m - n] = make_float4(rx, ry, rz, -1.0f);
        start[m - n] = 0;
        count[m - n] = -1;
        arrived[m - n] = 0;  // level 0, no child cells yet, no reports
        sc->step = sc->step + 1;
    }
}

// ---- 2. tree build ------------------------------------------------------------
// buildtree.cl: concurrent insertion; a child slot is locked by CAS to -2 while a
// leaf is split, the finished sub-tree is published by a release store.  The
// tree *shape* is a function of the positions and the root box only; cell
// numbers depend on the allocation race (as in the reference).  Bodies are
// inserted in slot order, i.e. in the previous step's tree order: the loads of
// a lane's run are sequential and its consecutive bodies are spatial neighbours.
//
// B200 design, all of it about latency (the kernel issues < 25 % of its slots):
//  * Every lane owns a contiguous run of slots.  The lane remembers the path of its
//    previous body (cells never move or disappear during a build) and replays it
//    without touching memory -- the octant tests use centres recomputed with the
//    creation formula, same operands, same bits -- so only the last level or two
//    are dependent loads.  Lanes of a warp are a run apart: few lock conflicts.
//  * The warp meets once per round.  A lane that won a lock on an occupied leaf
//    first counts, in registers, how many cells separate the two bodies; all such
//    lanes then take their cells from ONE atomicSub on `bottom` (warp-aggregated,
//    exact and gap-free; one atomic per cell serialised ~0.48 N same-address
//    atomics, about 2 ms at N = 10^7) and build their whole chain in that round.
//    Indices still decrease in allocation order, so a child cell always has a
//    lower index than its parent, which sort relies on.
//  * Memory model: a new sub-tree is written with ordinary stores and published by
//    st.release.gpu into the locked slot (buildtree.cl:173-180); other threads
//    reach it only through that slot and read it with strong (ld.relaxed.gpu, L2)
//    loads whose addresses depend on the value loaded from the slot.
constexpr int kBuildThreads = 256;
constexpr int kPathCap = 24;  // remembered levels per lane (24 KB of shared memory per CTA); deeper levels are loaded

__global__ void __launch_bounds__(kBuildThreads) build_kernel(const float4 *__restrict__ body4, float4 *__restrict__ cell4,
                                                              int *child, int *__restrict__ start, int *__restrict__ count,
                                                              int *__restrict__ parent, int *__restrict__ arrived, Scalars *sc,
                                                              int n, int m) {
    constexpr unsigned kFull = 0xffffffffu;
    const float radius = sc->radius;
    const float4 root = cell4[m - n];
    const int lane = threadIdx.x & 31;
    const int threadsTotal = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int run = (n + threadsTotal - 1) / threadsTotal;  // bodies per lane
    int i = tid * run;
    const int iEnd = min(n, i + run);
    bool busy = i < iEnd;   // this lane still has bodies to insert
    bool fresh = true;      // the next round starts a new body
    int localMaxDepth = 1, spins = 0;
    // remembered path of the lane's previous body: the cells at depths 1..pathLen (shared memory, one column per
    // thread: conflict-free) and the body's position -- its octants are recomputed, not stored
    __shared__ int pathNode[kPathCap][kBuildThreads];
    int pathLen = 1;
    pathNode[0][threadIdx.x] = m;
    float ppx = 0.0f, ppy = 0.0f, ppz = 0.0f;
    int body = 0, node = m, depth = 1, path = 0;
    float4 p = root;
    float r = radius, cx = root.x, cy = root.y, cz = root.z;
    for (;;) {
        int need = 0;        // cells this lane must allocate in this round
        int oldBody = -1;
        int *slot = child;
        float4 q = root;
        if (busy) {
            if (fresh) {
                body = i;
                p = body4[body];
                node = m; depth = 1; r = radius;
                cx = root.x; cy = root.y; cz = root.z;
                path = octant(cx, cy, cz, p.x, p.y, p.z);
                // replay the remembered path while the new body takes the same octants as the previous one
                while (depth < pathLen && octant(cx, cy, cz, ppx, ppy, ppz) == path) {
                    const float ox = (path & 1) ? r : 0.0f, oy = (path & 2) ? r : 0.0f, oz = (path & 4) ? r : 0.0f;
                    r *= 0.5f;
                    cx = __fadd_rn(__fsub_rn(cx, r), ox);  // buildtree.cl:124-136
                    cy = __fadd_rn(__fsub_rn(cy, r), oy);
                    cz = __fadd_rn(__fsub_rn(cz, r), oz);
                    ++depth;
                    path = octant(cx, cy, cz, p.x, p.y, p.z);
                }
                node = pathNode[depth - 1][threadIdx.x];
                pathLen = depth;
                ppx = p.x; ppy = p.y; ppz = p.z;
                fresh = false;
            }
            slot = child + ((size_t)(node - n) * 8 + path);
            int ch = ld_relaxed(slot);
            while (ch >= n) {  // buildtree.cl:77-89: follow the path to a leaf slot
                if (depth < kPathCap) pathNode[depth][threadIdx.x] = ch;
                node = ch;
                ++depth;
                const float ox = (path & 1) ? r : 0.0f, oy = (path & 2) ? r : 0.0f, oz = (path & 4) ? r : 0.0f;
                r *= 0.5f;
                cx = __fadd_rn(__fsub_rn(cx, r), ox);
                cy = __fadd_rn(__fsub_rn(cy, r), oy);
                cz = __fadd_rn(__fsub_rn(cz, r), oz);
                path = octant(cx, cy, cz, p.x, p.y, p.z);
                slot = child + ((size_t)(node - n) * 8 + path);
                ch = ld_relaxed(slot);
            }
            pathLen = min(depth, kPathCap);
            if (ch != kLock && atomicCAS(slot, ch, kLock) == ch) {
                if (ch == -1) {
                    st_relaxed(slot, body);  // buildtree.cl:98-101
                    localMaxDepth = max(localMaxDepth, depth);
                    fresh = true;
                    busy = ++i < iEnd;
                } else {  // buildtree.cl:102-180: count the cells that separate the two bodies
                    oldBody = ch;
                    q = body4[ch];
                    float tr = r, tx = cx, ty = cy, tz = cz;
                    int tp = path;
                    for (;;) {
                        ++need;
                        const float ox = (tp & 1) ? tr : 0.0f, oy = (tp & 2) ? tr : 0.0f, oz = (tp & 4) ? tr : 0.0f;
                        tr *= 0.5f;
                        tx = __fadd_rn(__fsub_rn(tx, tr), ox);
                        ty = __fadd_rn(__fsub_rn(ty, tr), oy);
                        tz = __fadd_rn(__fsub_rn(tz, tr), oz);
                        tp = octant(tx, ty, tz, p.x, p.y, p.z);
                        if (tp != octant(tx, ty, tz, q.x, q.y, q.z) || depth + need > kMaxDepth) break;
                    }
                }
            } else if (ch == kLock && (++spins & 63) == 0) {
                if (*reinterpret_cast<volatile int *>(&sc->error) != 0) busy = false;
                if (spins > kSpinBudget) { atomicCAS(&sc->error, 0, 2); busy = false; }
            }
        }
        // ---- the warp meets: aggregated cell allocation -------------------------------------------------
        if (__any_sync(kFull, need != 0)) {
            int incl = need;  // inclusive prefix sum over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(kFull, incl, 31);
            int top = 0;
            if (lane == 0) top = atomicSub(&sc->bottom, total);  // buildtree.cl:109, once for the warp
            top = __shfl_sync(kFull, top, 0);
            if (need != 0) {
                int cell = top - 1 - (incl - need);  // this lane's cells: cell, cell-1, ..., cell-need+1
                if (cell - need + 1 <= n || depth + need > kMaxDepth) {  // buildtree.cl:112-119 / calculateforce.cl:69-73
                    sc->error = 1;
                    __threadfence();
                    st_relaxed(slot, oldBody);  // give the leaf back so that nobody spins on it
                    busy = false;
                } else {
                    const int patch = cell;
                    const int node0 = node;  // the cell whose leaf slot is being split
                    int cur = node, curPath = path;
                    for (;;) {
                        ++depth;
                        const float ox = (curPath & 1) ? r : 0.0f;  // buildtree.cl:124-126
                        const float oy = (curPath & 2) ? r : 0.0f;
                        const float oz = (curPath & 4) ? r : 0.0f;
                        r *= 0.5f;
                        cx = __fadd_rn(__fsub_rn(cx, r), ox);  // buildtree.cl:134-136, left to right
                        cy = __fadd_rn(__fsub_rn(cy, r), oy);
                        cz = __fadd_rn(__fsub_rn(cz, r), oz);
                        cell4[cell - n] = make_float4(cx, cy, cz, -1.0f);
                        start[cell - n] = -1;
                        count[cell - n] = -1;
                        parent[cell - n] = cur;  // for the counter-driven summarise
                        const int qPath = octant(cx, cy, cz, q.x, q.y, q.z);
                        const int pPath = octant(cx, cy, cz, p.x, p.y, p.z);
                        // level << 16 | #child cells << 8 | reports received; the new cell's level is depth - 1 (root = 0)
                        arrived[cell - n] = ((depth - 1) << 16) | ((qPath == pPath) ? (1 << 8) : 0);
                        int *row = child + (size_t)(cell - n) * 8;
                        const int4 empty = make_int4(-1, -1, -1, -1);
                        reinterpret_cast<int4 *>(row)[0] = empty;
                        reinterpret_cast<int4 *>(row)[1] = empty;
                        if (cell != patch) child[(size_t)(cur - n) * 8 + curPath] = cell;  // :141-146
                        if (depth <= kPathCap) pathNode[depth - 1][threadIdx.x] = cell;  // the new cell joins the remembered path
                        cur = cell;
                        curPath = pPath;
                        if (qPath != pPath) {
                            row[qPath] = oldBody;  // :152
                            row[pPath] = body;     // :169
                            break;
                        }
                        --cell;
                    }
                    pathLen = min(depth, kPathCap);
                    node = cur;
                    path = curPath;
                    atomicAdd(arrived + (node0 - n), 1 << 8);  // the leaf's cell gains a child cell
                    st_release(slot, patch);  // :173-180 publish the sub-tree by replacing the lock
                    localMaxDepth = max(localMaxDepth, depth);
                    fresh = true;
                    busy = ++i < iEnd;
                }
            }
        }
        if (!__any_sync(kFull, busy)) break;
    }
    for (int o = 16; o > 0; o >>= 1) localMaxDepth = max(localMaxDepth, __shfl_xor_sync(kFull, localMaxDepth, o));
    if (lane == 0 && localMaxDepth > 1) atomicMax(&sc->maxDepth, localMaxDepth);  // :199
}

// ---- 3. summarise ---------------------------------------------------------------
// summarizetree.cl: bottom-up centre of mass, body counts, child compaction.
// The reference (and Burtscher's original) walks the cells in ascending index
// order and spins on children that are not ready; on 10^7 bodies most threads
// then sit in long parent-child chains.  Here the pass is counter driven and
// never waits: every cell collects one report per child cell plus one from its
// own thread (`arrived`: build keeps the number of child cells in bits 8-15,
// reports count up in bits 0-7; a small array that stays in L2); whoever
// reports last summarises the cell, reports to the parent and climbs on.  A
// report is one atom.add.acq_rel.gpu: release for the record just written,
// acquire for the siblings' records.  Children are summed in octant order,
// which makes the result independent of timing and bit-identical to the oracle.
// The summarising thread also writes the force walk's record of the cell,
// including the opening threshold of its children (calculateforce.cl:52-67:
// dq[level] = radius^2 * 0.25^level / THETA + EPSILON, the same iteration).
constexpr int kSummThreads = 256;

__device__ __forceinline__ void fill_dq(float *dq, const Scalars *sc, float thetaMacro, float eps) {
    // calculateforce.cl:52-67
    const float radius = sc->radius;
    const int maxDepth = min(sc->maxDepth, kMaxDepth);
    float v = __fmul_rn(radius, radius);
    if (thetaMacro > 0.0f) v = __fdiv_rn(v, thetaMacro);
    for (int i = 0; i < maxDepth; ++i) {
        dq[i] = __fadd_rn(v, eps);
        v = __fmul_rn(0.25f, v);
    }
}

__global__ void __launch_bounds__(kSummThreads, 4) summarize_kernel(const float4 *__restrict__ body4, float4 *__restrict__ cell4,
                                                                 int *__restrict__ child, float4 *__restrict__ octet,
                                                                 int2 *__restrict__ ometa, int *__restrict__ meta,
                                                                 int *__restrict__ count, const int *__restrict__ parent,
                                                                 int *arrived, Scalars *sc, int n, int m, float thetaMacro,
                                                                 float eps) {
    __shared__ float dq[kMaxDepth];
    if (sc->error != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->bottom = m;  // buildtree.cl:117
        return;
    }
    if (threadIdx.x == 0) fill_dq(dq, sc, thetaMacro, eps);
    __syncthreads();
    const int bottom = sc->bottom;
    const int stride = gridDim.x * blockDim.x;
    for (int first = bottom + blockIdx.x * blockDim.x + threadIdx.x; first <= m; first += stride) {
        int cell = first;
        // A cell is summarised by whoever reports last among its child cells and its own thread.
        int old = atom_add_acquire(arrived + (cell - n), 1);
        if ((old & 0xff) != ((old >> 8) & 0xff)) continue;
        for (;;) {
            const int level = min(old >> 16, kMaxDepth - 1);
            int *row = child + (size_t)(cell - n) * 8;
            const int4 lo = __ldcg(reinterpret_cast<const int4 *>(row)), hi = __ldcg(reinterpret_cast<const int4 *>(row) + 1);
            const int in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            int out[8];
            int used = 0, ncell = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = -1;
#pragma unroll
            for (int k = 0; k < 8; ++k)  // summarizetree.cl:77-81 compaction, octant order kept
                if (in[k] >= 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j == used) out[j] = in[k];
                    ++used;
                    ncell += in[k] >= n;
                }
            // all child records are final (child cells reported before this thread got here): fetch them together
            float4 c[8];
            int cnt[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                c[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                cnt[k] = 0;
                if (out[k] >= n) {
                    c[k] = __ldcg(cell4 + (out[k] - n));
                    cnt[k] = __ldcg(count + (out[k] - n));
                } else if (out[k] >= 0) {
                    c[k] = body4[out[k]];
                    cnt[k] = 1;
                }
            }
            float cm = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
            int bodies = 0;  // summarizetree.cl:98-105,118
            float4 *orow = octet + (size_t)(cell - n) * 8;
            int2 *mrow = ometa + (size_t)(cell - n) * 8;
            const int thrBits = __float_as_int(dq[level]);
            const int bodyThr = __float_as_int(-1.0f);
            int cpos = 0, bpos = ncell;  // the force walk's record: child cells first, then child bodies
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (out[k] < 0) break;
                bodies += cnt[k] & kCountMask;
                cm = __fadd_rn(cm, c[k].w);  // summarizetree.cl:107-110, octant order
                cx = fmaf(c[k].x, c[k].w, cx);
                cy = fmaf(c[k].y, c[k].w, cy);
                cz = fmaf(c[k].z, c[k].w, cz);
                if (out[k] >= n) {
                    mrow[cpos] = make_int2(thrBits, (out[k] - n) | ((cnt[k] >> 28) << 27));
                    orow[cpos++] = c[k];
                } else {
                    mrow[bpos] = make_int2(bodyThr, -1);
                    orow[bpos++] = c[k];
                }
            }
            // whole 32-byte sectors only: a partially written sector costs a DRAM read (merge) when it leaves L2
            // (measured: 0.21 GB less DRAM read traffic, 0.996 -> 0.922 ms at 10^7 bodies)
            if (used & 1) orow[used] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 1; k < 8; ++k)
                if (k >= used && (k & 3) != 0 && (k & ~3) < used) mrow[k] = make_int2(0, 0);
            reinterpret_cast<int4 *>(row)[0] = make_int4(out[0], out[1], out[2], out[3]);
            reinterpret_cast<int4 *>(row)[1] = make_int4(out[4], out[5], out[6], out[7]);
            meta[cell - n] = ncell | ((used - ncell) << 4);  // deep walk: #child cells, #child bodies
            const float inv = __frcp_rn(cm);  // summarizetree.cl:161: 1.0f / cellMass, correctly rounded
            __stcg(cell4 + (cell - n), make_float4(__fmul_rn(cx, inv), __fmul_rn(cy, inv), __fmul_rn(cz, inv), cm));
            __stcg(count + (cell - n), bodies | ((used - 1) << 28));  // summarizetree.cl:160 (+ #children for the walk entry)
            if (cell == m) {  // the root
                sc->rootEntry = (m - n) | ((used - 1) << 27);
                break;
            }
            // report to the parent; the child that reports last continues with it (summarizetree.cl:170: release
            // this cell's record, acquire the siblings')
            const int par = parent[cell - n];
            old = atom_add_acq_rel(arrived + (par - n), 1);
            if ((old & 0xff) != ((old >> 8) & 0xff)) break;  // not the last report: somebody else continues
            cell = par;
        }
    }
}

// ---- 4. sort ----------------------------------------------------------------------
// sort.cl: top-down propagation of `start`, bodies written in DFS order.  One
// thread per cell, descending index (parents first), waits for start >= 0.
// Launched cooperatively (all CTAs co-resident, or the launch waits): a waiting
// thread's parent is always being processed by a resident thread.
constexpr int kSortThreads = 256;

__global__ void __launch_bounds__(kSortThreads) sort_kernel(const int *__restrict__ child, const int *__restrict__ count,
                                                            int *start, int *__restrict__ perm, Scalars *sc, int n, int m) {
    if (sc->error != 0) return;
    const int bottom = sc->bottom;
    const int stride = gridDim.x * blockDim.x;
    for (int cell = m - (blockIdx.x * blockDim.x + threadIdx.x); cell >= bottom; cell -= stride) {
        int s, spins = 0;
        while ((s = ld_relaxed(start + (cell - n))) < 0) {  // sort.cl:36-39 (start is the only datum passed: no fence needed)
            if ((++spins & 255) == 0 && (spins > kSpinBudget || *reinterpret_cast<volatile int *>(&sc->error) != 0)) {
                atomicCAS(&sc->error, 0, 2);
                return;
            }
        }
        const int *row = child + (size_t)(cell - n) * 8;
        const int4 lo = reinterpret_cast<const int4 *>(row)[0], hi = reinterpret_cast<const int4 *>(row)[1];
        const int ch[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (ch[k] < 0) break;  // compacted
            if (ch[k] >= n) {      // sort.cl:44-56
                st_relaxed(start + (ch[k] - n), s);
                s += count[ch[k] - n] & kCountMask;
            } else {               // sort.cl:59-65
                perm[s++] = ch[k];
            }
        }
    }
}

// ---- 5. force -------------------------------------------------------------------
// calculateforce.cl: a vote group of 16 consecutive sorted bodies walks the tree
// together; a cell is used as a point mass only if *all* bodies of the group
// are far enough (work_group_all, :145), bodies are always used.  The set of
// (group, node) interactions is exactly the reference's; the order in which a
// body sums them differs, which moves the fp32 sum by rounding only.
//
// Both kernels are built around Blackwell's packed fp32 pipe (FADD2 / FMUL2 /
// FFMA2: two IEEE fp32 operations per issue slot): every lane carries TWO
// consecutive sorted bodies, a warp 64 bodies = four 16-body vote groups (lanes
// 8g..8g+7 are group g).  Results go to the tree-order acceleration buffer(s)
// `dst` (the rank's own and, in a multi-GPU run with peer memory, every other
// rank's over NVLink: the all-gather is fused into the walk's epilogue).

__device__ __forceinline__ float rsqrt_fast(float x) {
    // r^2 >= EPSILON > 0 is never subnormal: the bare MUFU.RSQ (2 ulp, calculateforce.cl:146 allows rsqrt's 2 ulp)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// shared memory through explicit 32-bit shared-window addresses
__device__ __forceinline__ void lds_v2(unsigned a, int &x, int &y) { asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a) : "memory"); }
__device__ __forceinline__ void sts_v2(unsigned a, int x, int y) { asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ float4 lds_v4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v4(unsigned a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int lds_s32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_s32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
// global -> shared without a register round trip (LDGSTS)
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ const char *lane_address(const char *base, int rel, int stride) {
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(rel), "r"(stride), "l"(base));
    return reinterpret_cast<const char *>(a);
}

// calculateforce.cl:146-151 for the lane's two bodies; mw = child mass, or 0 for a lane whose group does not use the child
__device__ __forceinline__ void force_accumulate(float2 dx, float2 dy, float2 dz, float2 r2, float mw, float2 &ax, float2 &ay,
                                                 float2 &az) {
    const float2 rinv = make_float2(rsqrt_fast(r2.x), rsqrt_fast(r2.y));
    const float2 f = __fmul2_rn(__fmul2_rn(__fmul2_rn(make_float2(mw, mw), rinv), rinv), rinv);
    ax = __ffma2_rn(dx, f, ax);
    ay = __ffma2_rn(dy, f, ay);
    az = __ffma2_rn(dz, f, az);
}

#define BH_DIST(c, S)                                                                                                  \
    const float2 dx##S = __fadd2_rn(make_float2((c).x, (c).x), npx); /* c - p, exactly */                              \
    const float2 dy##S = __fadd2_rn(make_float2((c).y, (c).y), npy);                                                   \
    const float2 dz##S = __fadd2_rn(make_float2((c).z, (c).z), npz);                                                   \
    const float2 r2##S = __fadd2_rn(__ffma2_rn(dz##S, dz##S, __ffma2_rn(dy##S, dy##S, __fmul2_rn(dx##S, dx##S))), eps2); /* :138-143 */

// Destinations of a slice's accelerations: the rank's own tree-order buffer and, in a multi-GPU run with
// peer memory, the same buffer of every other rank mapped over NVLink (CUDA IPC).  `phaseStride` (float4s)
// selects the half used in this step (step & 1): peers may already store the next step's slices while this
// step's are still being read.  0 = single buffer.
constexpr int kMaxPeers = 16;
struct PeerBuffers {
    float4 *buf[kMaxPeers];
    int count;
    unsigned phaseStride;
};

// ---- 5a. the walk: one private work list per vote group ---------------------------------------------------------
// A lane carries FOUR consecutive sorted bodies (two packed pairs), four lanes are one 16-body vote group, a
// warp works on eight groups at a time.  Every group keeps, in shared memory, a stack of cells it has decided
// to open and a LIFO of child records waiting to be tested.  A warp-wide step tests ONE queued child per group
// -- eight different children, each against its own group's 16 bodies (LDS.128 with one address per group) --
// so no lane ever works on a node its group does not need (a shared walk of the groups' union costs ~25 % more
// child tests), and the per-cell bookkeeping of a stack walk (pop, row fetch, loop control: a third of the
// issue slots of a pop-per-cell design) is paid once per pass of kWalkBatch children: at the top of a pass
// every group pops up to kWalkTrips cells from its stack -- while its queue has room for any cell -- and
// LDGSTS copies their walk records (children {x,y,z,m} and {threshold, entry}: one 128-byte and one 64-byte
// line per cell) straight into the queue.  What a group does depends on its own state only, so its result is
// bit-identical in whatever slot, warp, grid or slice it is computed (single-GPU run == any rank's run).
// Per child (128 interactions): LDS.128 + LDS.32, 14 packed fp32 (distances), FMNMX + FMNMX3 + FSETP + VOTE +
// LOP3 (the group's vote, :145), 4 MUFU.RSQ, 12 packed fp32 (the interactions; zero mass if the group opens
// the cell; the mass is multiplied in last so that the vote is off the critical path), predicated push.
// CTAs are persistent: a group slot whose walk is finished writes its accelerations and takes the next group of
// its warp's chunk (eight consecutive groups, drawn from a global ticket), so a warp never waits for its slowest
// group and all warps of the grid finish within one group walk of each other.
// A group's cell stack holds kWalkSCap entries in shared memory; in deep trees (7 open siblings per level)
// its bottom entries are spilled to a per-slot global buffer and come back when the shared part runs dry, so
// any tree the reference accepts (64 levels) is walked by this kernel.
constexpr int kWalkThreads = 128;
constexpr int kWalkWarps = kWalkThreads / 32;
constexpr int kWalkGroups = 8;        // vote groups per warp (4 lanes x 4 bodies)
#ifndef BH_WALK_BATCH  // (tuning knobs: scripts/walk_variants.sh builds and times alternatives)
#define BH_WALK_BATCH 16
#define BH_WALK_SUB 8
#define BH_WALK_TRIPS 5
#define BH_WALK_SCAP 96
#define BH_WALK_CTAS 4
#endif
constexpr int kWalkBatch = BH_WALK_BATCH;  // children tested per group and pass
constexpr int kWalkSub = BH_WALK_SUB;      // ... in sub-batches of this many, whose loads are issued up front
constexpr int kWalkTrips = BH_WALK_TRIPS;  // cells a group pops per pass, at most
constexpr int kWalkQCap = kWalkBatch + 8;  // queued children per group: a group pops while 8 slots (any cell) are free
constexpr int kWalkSlots = kWalkBatch + kWalkQCap;  // the lowest kWalkBatch slots hold zero-mass dummies
constexpr int kWalkSCap = BH_WALK_SCAP;    // stacked cells per group in shared memory
constexpr int kWalkSpill = 64;        // entries moved to / from the global spill buffer at a time
constexpr int kWalkSpillCap = 1024;   // spill entries per group slot (7 * 64 + 8 would do)
constexpr int kWalkCtasPerSM = BH_WALK_CTAS;
static_assert(kWalkBatch % kWalkSub == 0 && kWalkSCap >= kWalkSpill + kWalkBatch + 8, "walk tuning");

struct WalkShared {
    // +1 padding: the groups' arrays start 4 (2, 1) banks apart, so equal indices in different groups do not collide
    int stk[kWalkWarps][kWalkGroups][kWalkSCap + 1];
    float4 rec[kWalkWarps][kWalkGroups][kWalkSlots + 1];
    int2 meta[kWalkWarps][kWalkGroups][kWalkSlots + 1];
};

template <bool COUNT, bool POT>
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSM) walk_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ octet,
                                                                            const int2 *__restrict__ ometa, const int *__restrict__ perm,
                                                                            const PeerBuffers dst, Scalars *sc, int *__restrict__ spill, int n,
                                                                            int first, int cnt, float eps) {
    extern __shared__ float4 walkSharedRaw[];  // dynamic: more than the 48 KB a static array may have
    WalkShared &sh = *reinterpret_cast<WalkShared *>(walkSharedRaw);
    if (sc->error != 0) return;
    if (sc->maxDepth > kMaxDepth) {  // calculateforce.cl:69-73
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->error = 1;
        return;
    }
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, l = lane & 3;
    if (l == 0) {
#pragma unroll
        for (int i = 0; i < kWalkBatch; ++i) {  // dummies: zero mass, always accepted, never counted
            sh.rec[warp][g][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            sh.meta[warp][g][i] = make_int2(__float_as_int(-1.0f), -2);
        }
    }
    __syncthreads();
    const int end = first + cnt;
    const int totalGroups = (cnt + 15) >> 4;
    const int rootEntry = sc->rootEntry;
    const size_t phase = (size_t)(sc->step & 1) * dst.phaseStride;
    unsigned gm = 0xfu << (lane & ~3);  // lanes of my group
    asm volatile("" : "+r"(gm));        // keep it in a register (ptxas would rematerialise it per vote)
    const float2 eps2 = make_float2(eps, eps);
    // shared memory through 32-bit shared-window addresses kept in registers
    const unsigned stkBase = (unsigned)__cvta_generic_to_shared(&sh.stk[warp][g][0]);
    const unsigned stkLimit = stkBase + 4u * (kWalkSCap - kWalkBatch);  // a batch pushes at most kWalkBatch cells
    const unsigned qBase = (unsigned)__cvta_generic_to_shared(&sh.rec[warp][g][kWalkBatch]);
    const unsigned mBase = (unsigned)__cvta_generic_to_shared(&sh.meta[warp][g][kWalkBatch]);
    unsigned stkTop = stkBase, qTop = qBase, mTop = mBase;  // one past the top entry; the same in all lanes of a group
    int *const mySpill = spill + ((size_t)(blockIdx.x * kWalkWarps + warp) * kWalkGroups + g) * kWalkSpillCap;
    int spilled = 0;  // entries of my group's stack that live in the spill buffer (below the shared-memory part)
    // the group this slot works on; the warp draws chunks of eight consecutive groups from a global ticket, so that
    // all warps of the grid finish within one group walk of each other whatever the groups cost
    bool active = false;               // the slot has a group
    bool fetch = true;                 // the slot wants its first group
    int chunkNext = 0, chunkEnd = 0;   // unassigned groups of the warp's current chunk (the same in all lanes)
    int k0 = 0, nact = 0;
    float2 npxA = make_float2(0.f, 0.f), npyA = npxA, npzA = npxA, npxB = npxA, npyB = npxA, npzB = npxA;
    float2 axA = npxA, ayA = npxA, azA = npxA, axB = npxA, ayB = npxA, azB = npxA;
    float2 potA = npxA, potB = npxA;  // POT: sum of m / sqrt(r^2 + eps) over the nodes used (the tree's potential, self term included)
    unsigned long long nInter = 0, nOpen = 0;
    for (;;) {
        // ---- group slots: finished walks write their accelerations and take the next group ---------------------
        const bool finished = active && stkTop == stkBase && qTop == qBase && spilled == 0;
        if (__any_sync(kFull, finished || fetch)) {
            if (finished) {
                const float4 a[4] = {make_float4(axA.x, ayA.x, azA.x, potA.x), make_float4(axA.y, ayA.y, azA.y, potA.y),
                                     make_float4(axB.x, ayB.x, azB.x, potB.x), make_float4(axB.y, ayB.y, azB.y, potB.y)};
                for (int r = 0; r < dst.count; ++r) {  // own buffer, then the peers' (NVLink stores)
                    float4 *out = dst.buf[r] + phase + k0;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t < nact) out[t] = a[t];
                }
            }
            const bool want = finished || fetch;
            const unsigned wantMask = __ballot_sync(kFull, want && l == 0);  // one bit per slot that wants a group
            const int need = __popc(wantMask), have = chunkEnd - chunkNext;
            int fresh = 0;
            if (need > have) {  // the rest of the current chunk does not cover them: draw the next chunk
                if (lane == 0) fresh = atomicAdd(&sc->walkTicket, 1) * kWalkGroups;
                fresh = __shfl_sync(kFull, fresh, 0);
            }
            if (want) {
                fetch = false;
                const int r = __popc(wantMask & ((1u << (lane & ~3)) - 1u));  // my rank among the slots that want one
                const int grp = r < have ? chunkNext + r : fresh + (r - have);
                active = grp < totalGroups;
                nact = 0;
                if (active) {
                    const int gfirst = first + grp * 16;  // the group's first sorted slot (exists)
                    k0 = gfirst + 4 * l;
                    nact = min(4, max(0, end - k0));
                    // a slot past the end borrows the position of the group's first body: its vote then equals that body's
                    float4 p[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int s = t < nact ? k0 + t : gfirst;
                        p[t] = body4[perm ? perm[s] : s];  // perm == nullptr: bodies lie in tree order
                    }
                    npxA = make_float2(-p[0].x, -p[1].x); npyA = make_float2(-p[0].y, -p[1].y); npzA = make_float2(-p[0].z, -p[1].z);
                    npxB = make_float2(-p[2].x, -p[3].x); npyB = make_float2(-p[2].y, -p[3].y); npzB = make_float2(-p[2].z, -p[3].z);
                    axA = ayA = azA = axB = ayB = azB = potA = potB = make_float2(0.f, 0.f);
                    sts_s32(stkBase, rootEntry);
                    stkTop = stkBase + 4;
                    qTop = qBase;
                    mTop = mBase;
                }
            }
            if (need > have) { chunkNext = fresh + (need - have); chunkEnd = fresh + kWalkGroups; } else { chunkNext += need; }
            if (__all_sync(kFull, !active)) break;
        }
        // ---- refill: every pass, every group pops up to kWalkTrips cells while its queue has room for any cell ------
        // (a group's schedule depends on its own state only: its result is bit-reproducible whichever slot, warp or
        // launch configuration -- whole universe or one rank's slice -- it runs in)
        if (__any_sync(kFull, stkTop == stkBase && spilled != 0)) {  // (cold) a dry stack takes spilled entries back
            if (stkTop == stkBase && spilled != 0) {
                const int take = min(spilled, kWalkSpill);
                spilled -= take;
                for (int t = l; t < take; t += 4) sts_s32(stkBase + 4u * (unsigned)t, mySpill[spilled + t]);
                stkTop = stkBase + 4u * (unsigned)take;
            }
            __syncwarp();
        }
        {
            int room = kWalkQCap - ((int)(qTop - qBase) >> 4);
#pragma unroll
            for (int r = 0; r < kWalkTrips; ++r) {
                if (active && stkTop != stkBase && room >= 8) {
                    const int e = lds_s32(stkTop - 4u);
                    const int c = ((e >> 27) & 7) + 1;  // children of the cell on top of the stack
                    const int rel = e & kEntryMask;
                    const float4 *src = octet + (size_t)rel * 8 + l;
                    const int2 *msrc = ometa + (size_t)rel * 8 + l;
                    const unsigned d = qTop + 16u * (unsigned)l, md = mTop + 8u * (unsigned)l;
                    if (l < c) { cp_async16(d, src); cp_async8(md, msrc); }
                    if (l + 4 < c) { cp_async16(d + 64u, src + 4); cp_async8(md + 32u, msrc + 4); }
                    stkTop -= 4u;
                    qTop += 16u * (unsigned)c;
                    mTop += 8u * (unsigned)c;
                    room -= c;
                }
            }
            cp_async_wait_all();
            __syncwarp();
        }
        if (__any_sync(kFull, stkTop > stkLimit)) {  // (cold) no room for the pushes of another batch: spill the bottom entries
            if (stkTop > stkLimit) {
                if (spilled + kWalkSpill > kWalkSpillCap) {
                    atomicCAS(&sc->error, 0, 2);  // cannot happen for trees of at most 64 levels
                } else {
                    const int rest = ((int)(stkTop - stkBase) >> 2) - kWalkSpill;  // entries that stay
                    int keep[kWalkSpill / 4];
#pragma unroll
                    for (int t = 0; t < kWalkSpill / 4; ++t) {
                        mySpill[spilled + l + 4 * t] = lds_s32(stkBase + 4u * (unsigned)(l + 4 * t));
                        keep[t] = (l + 4 * t < rest) ? lds_s32(stkBase + 4u * (unsigned)(kWalkSpill + l + 4 * t)) : 0;
                    }
                    __syncwarp(gm);
#pragma unroll
                    for (int t = 0; t < kWalkSpill / 4; ++t)
                        if (l + 4 * t < rest) sts_s32(stkBase + 4u * (unsigned)(l + 4 * t), keep[t]);
                    spilled += kWalkSpill;
                    stkTop -= 4u * kWalkSpill;
                    if (l == 0) atomicAdd(&sc->walkSpills, 1);
                }
            }
            __syncwarp();
        }
        // ---- one batch: the top kWalkBatch queued children of every group (dummies below a short queue) ------------
#pragma unroll
        for (int h = 0; h < kWalkBatch; h += kWalkSub) {
            float4 cs[kWalkSub];
            float thrs[kWalkSub];
#pragma unroll
            for (int i = 0; i < kWalkSub; ++i) {  // the loads of a sub-batch first: their latency overlaps the first children's math
                cs[i] = lds_v4(qTop - 16u * (unsigned)(h + i + 1));
                thrs[i] = lds_f32(mTop - 8u * (unsigned)(h + i + 1));
            }
#pragma unroll
            for (int i = 0; i < kWalkSub; ++i) {
                const float4 c = cs[i];
                const float thr = thrs[i];
                const unsigned entAddr = mTop - 8u * (unsigned)(h + i + 1) + 4u;
                const float2 cx2 = make_float2(c.x, c.x), cy2 = make_float2(c.y, c.y), cz2 = make_float2(c.z, c.z);
                const float2 dxA = __fadd2_rn(cx2, npxA), dyA = __fadd2_rn(cy2, npyA), dzA = __fadd2_rn(cz2, npzA);  // c - p, exactly
                const float2 dxB = __fadd2_rn(cx2, npxB), dyB = __fadd2_rn(cy2, npyB), dzB = __fadd2_rn(cz2, npzB);
                const float2 r2A = __fadd2_rn(__ffma2_rn(dzA, dzA, __ffma2_rn(dyA, dyA, __fmul2_rn(dxA, dxA))), eps2);  // :138-143
                const float2 r2B = __fadd2_rn(__ffma2_rn(dzB, dzB, __ffma2_rn(dyB, dyB, __fmul2_rn(dxB, dxB))), eps2);
                const float rmin = fminf(fminf(r2A.x, r2A.y), fminf(r2B.x, r2B.y));
                const bool far = !(rmin < thr);  // r^2 >= dq for the lane's four bodies (a NaN never opens: no entry of a body is ever pushed)
                const unsigned nearMask = __ballot_sync(kFull, !far);
                const bool open = (nearMask & gm) != 0u;  // :145 work_group_all failed: my group opens the cell
                const float mw = open ? 0.0f : c.w;
                const float2 mw2 = make_float2(mw, mw);
                const float2 rinvA = make_float2(rsqrt_fast(r2A.x), rsqrt_fast(r2A.y)), rinvB = make_float2(rsqrt_fast(r2B.x), rsqrt_fast(r2B.y));
                // :146-148 m r^-3; the mass (the only operand that waits for the vote) is multiplied in last
                const float2 fA = __fmul2_rn(__fmul2_rn(__fmul2_rn(rinvA, rinvA), rinvA), mw2);
                const float2 fB = __fmul2_rn(__fmul2_rn(__fmul2_rn(rinvB, rinvB), rinvB), mw2);
                axA = __ffma2_rn(dxA, fA, axA); ayA = __ffma2_rn(dyA, fA, ayA); azA = __ffma2_rn(dzA, fA, azA);  // :149-151
                axB = __ffma2_rn(dxB, fB, axB); ayB = __ffma2_rn(dyB, fB, ayB); azB = __ffma2_rn(dzB, fB, azB);
                if (POT) {
                    potA = __ffma2_rn(mw2, rinvA, potA);
                    potB = __ffma2_rn(mw2, rinvB, potB);
                }
                if (open) {  // all lanes of the group store the same entry to the same address
                    sts_s32(stkTop, lds_s32(entAddr));
                    stkTop += 4;
                }
                if (COUNT) {
                    if (open) nOpen += nact;
                    else if (lds_s32(entAddr) != -2) nInter += nact;
                }
            }
        }
        qTop = max(qTop - 16u * kWalkBatch, qBase);
        mTop = max(mTop - 8u * kWalkBatch, mBase);
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            nInter += __shfl_xor_sync(kFull, nInter, o);
            nOpen += __shfl_xor_sync(kFull, nOpen, o);
        }
        if (lane == 0) {
            atomicAdd(&sc->interactions, nInter);
            atomicAdd(&sc->opens, nOpen);
        }
    }
}

// ---- 5b. the deep walk: one shared stack per warp, 64 levels ---------------------------------------------------
// Round 1's walk, kept for the non-reference 32-wide vote and as a second implementation to validate and time
// walk_kernel against (bh_set_force_deep_walk).  A stack entry is {cell - N, one bit per group that still needs the
// cell | depth << 1}; a group that accepted a cell is simply absent from the mask of its children; the warp
// walks the union of its groups' trees.  Per pop one LDG.128 per lane brings the cell's walk record (8
// children + 8 {threshold, entry} pairs) into a per-warp shared-memory row; the children are consumed with
// broadcast LDS.128, child cells first and in pairs (two independent distance chains, one VOTE.ALL for both
// when every body of the warp is far from both).
constexpr int kStackCap = 7 * kMaxDepth + 8;  // a popped cell pushes at most 8 children, 7 stay while the 8th is walked
constexpr int kForce2Threads = 128;
constexpr int kForce2Bodies = 2 * kForce2Threads;  // per CTA

template <int VOTE, bool COUNT>
__global__ void __launch_bounds__(kForce2Threads) deep_walk_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ octet,
                                                                    const int2 *__restrict__ ometa, const int *__restrict__ meta,
                                                                    const int *__restrict__ perm, const PeerBuffers dst, Scalars *sc,
                                                                    int n, int m, int first, int cnt, float thetaMacro, float eps) {
    __shared__ float dq[kMaxDepth];
    __shared__ int2 stack[kForce2Threads / 32][kStackCap];  // {cell - N, group bits | depth << 1}
    __shared__ float4 stage[kForce2Threads / 32][12];       // the popped cell's walk record: 8 children, 8 {threshold, entry}
    if (sc->error != 0) return;
    const int maxDepth = sc->maxDepth;
    if (maxDepth > kMaxDepth) {  // calculateforce.cl:69-73
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->error = 1;
        return;
    }
    if (threadIdx.x == 0) fill_dq(dq, sc, thetaMacro, eps);
    __syncthreads();
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kLanesPerGroup = VOTE / 2;  // two bodies per lane
    constexpr unsigned kGroupLanes = (kLanesPerGroup == 32) ? kFull : ((1u << kLanesPerGroup) - 1u);
    constexpr unsigned kSpread = (kLanesPerGroup == 8) ? 0x01010101u : (kLanesPerGroup == 16) ? 0x00010001u : 1u;  // first lane of every group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int end = first + cnt;
    const unsigned stkBase = (unsigned)__cvta_generic_to_shared(stack[warp]);
    unsigned rowBase = (unsigned)__cvta_generic_to_shared(stage[warp]);
    unsigned dqBase = (unsigned)__cvta_generic_to_shared(dq);
    unsigned rowLane = rowBase + 16u * (unsigned)lane;
    // lanes 0-7 fetch the 8 child records, lanes 8-11 the 8 {threshold, entry} pairs
    const char *laneBase = lane < 8 ? reinterpret_cast<const char *>(octet + lane)
                                    : reinterpret_cast<const char *>(ometa) + 16 * ((lane - 8) & 3);
    int laneStride = lane < 8 ? 128 : 64;
    asm volatile("" : "+r"(rowBase), "+r"(dqBase), "+r"(rowLane), "+r"(laneStride));  // keep them in registers
    const size_t phase = (size_t)(sc->step & 1) * dst.phaseStride;
    const int chunks = (cnt + kForce2Bodies - 1) / kForce2Bodies;
    for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const int base = first + (chunk * (kForce2Threads / 32) + warp) * 64;
    // lane l carries sorted slots base+2l and base+2l+1 (same vote group)
    const int k0 = base + 2 * lane;
    const int nact = min(2, max(0, end - k0));  // bodies of this lane that exist
    const int gfirst = lane & ~(kLanesPerGroup - 1);  // first lane of my group
    unsigned gm = kGroupLanes << gfirst;              // lanes of my group
    unsigned gbit = 1u << gfirst;                     // my group's bit in a stack entry
    asm volatile("" : "+r"(gm), "+r"(gbit));          // keep both in registers (ptxas would rematerialise them per vote)
    // a slot past the end borrows the position of the group's first body: its vote then equals that body's
    const int kg = base + 2 * gfirst;
    const int s0 = (nact > 0) ? k0 : max(0, min(kg, end - 1)), s1 = (nact > 1) ? k0 + 1 : s0;
    const float4 p0 = body4[perm ? perm[s0] : s0], p1 = body4[perm ? perm[s1] : s1];
    const float2 npx = make_float2(-p0.x, -p1.x), npy = make_float2(-p0.y, -p1.y), npz = make_float2(-p0.z, -p1.z);
    const float2 eps2 = make_float2(eps, eps);
    float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
    unsigned long long nInter = 0, nOpen = 0;
    // groups with at least one existing body take part in the walk
    const unsigned lanesActive = __ballot_sync(kFull, nact > 0);
    unsigned startBits = 0;
#pragma unroll
    for (int g = 0; g < 32 / kLanesPerGroup; ++g)
        if (lanesActive & (kGroupLanes << (g * kLanesPerGroup))) startBits |= 1u << (g * kLanesPerGroup);
    unsigned sp = stkBase;  // address of the first free stack slot
    if (startBits != 0u) {
        sts_v2(sp, m - n, (int)startBits);  // depth 0
        sp += 8;
    }
    while (sp != stkBase) {
        sp -= 8;
        int rel, ey;
        lds_v2(sp, rel, ey);
        __syncwarp();  // every lane has read the entry and is done with the previous row
        if (lane < 12) sts_v4(rowLane, __ldg(reinterpret_cast<const float4 *>(lane_address(laneBase, rel, laneStride))));
        // REDUX puts the (warp-uniform) word into a uniform register: ptxas then knows that the
        // branches on it are uniform and emits no divergence guards around the votes
        const int mt = __reduce_or_sync(kFull, __ldg(meta + rel));
        const unsigned bits = (unsigned)ey & kSpread;
        const int dnext = (ey & 0x7e) + 2;  // (depth + 1) << 1
        const float thr = lds_f32(dqBase + ((unsigned)(ey & 0x7e) << 1));
        const bool mine = ((unsigned)ey & gbit) != 0u;
        const unsigned gmMine = mine ? gm : 0u;  // my group's lanes if my group still needs this cell
        const float mscale = mine ? 1.0f : 0.0f;
        const int ncell = mt & 15, nbody = mt >> 4;
        __syncwarp();
        // one child cell whose vote was not unanimous (or that was not tested as part of a pair)
#define BH_CELL_VOTE(c, S, far, j)                                                                                     \
    if (__all_sync(kFull, far)) { /* far enough for every body of the warp: every group that is here uses it */      \
        force_accumulate(dx##S, dy##S, dz##S, r2##S, __fmul_rn((c).w, mscale), ax, ay, az);                            \
        if (COUNT && mine) nInter += nact;                                                                             \
    } else {                                                                                                           \
        const unsigned near = __ballot_sync(kFull, !(far));                                                            \
        const unsigned open = __ballot_sync(kFull, (near & gmMine) != 0u);                                             \
        if (open) { /* :154-163 */                                                                                     \
            const int ch = lds_s32(rowBase + 132u + 8u * (j)) & kEntryMask;                                            \
            sts_v2(sp, ch, (int)((open & kSpread) | (unsigned)dnext));                                                 \
            sp += 8;                                                                                                   \
        }                                                                                                              \
        if (COUNT && (near & gmMine) != 0u) nOpen += nact;                                                             \
        if (bits & ~open) { /* at least one group uses the cell as a point mass */                                    \
            const bool use = mine && (near & gm) == 0u;                                                                \
            force_accumulate(dx##S, dy##S, dz##S, r2##S, use ? (c).w : 0.0f, ax, ay, az);                              \
            if (COUNT && use) nInter += nact;                                                                          \
        }                                                                                                              \
    }
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            if (j + 2 > ncell) break;
            const float4 c0 = lds_v4(rowBase + 16u * j), c1 = lds_v4(rowBase + 16u * (j + 1));
            BH_DIST(c0, 0)
            BH_DIST(c1, 1)
            const bool far0 = r20.x >= thr && r20.y >= thr, far1 = r21.x >= thr && r21.y >= thr;
            if (__all_sync(kFull, far0 && far1)) {
                const float mw0 = __fmul_rn(c0.w, mscale), mw1 = __fmul_rn(c1.w, mscale);
                force_accumulate(dx0, dy0, dz0, r20, mw0, ax, ay, az);
                force_accumulate(dx1, dy1, dz1, r21, mw1, ax, ay, az);
                if (COUNT && mine) nInter += 2 * nact;
            } else {
                BH_CELL_VOTE(c0, 0, far0, j)
                BH_CELL_VOTE(c1, 1, far1, j + 1)
            }
        }
        if (ncell & 1) {  // the odd one
            const int j = ncell - 1;
            const float4 c0 = lds_v4(rowBase + 16u * (unsigned)j);
            BH_DIST(c0, 0)
            const bool far0 = r20.x >= thr && r20.y >= thr;
            BH_CELL_VOTE(c0, 0, far0, j)
        }
        const unsigned brow = rowBase + 16u * (unsigned)ncell;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {  // then child bodies, also two at a time: always used (:145 child < NBODIES)
            if (j + 2 > nbody) break;
            const float4 c0 = lds_v4(brow + 16u * j), c1 = lds_v4(brow + 16u * (j + 1));
            BH_DIST(c0, 0)
            BH_DIST(c1, 1)
            force_accumulate(dx0, dy0, dz0, r20, __fmul_rn(c0.w, mscale), ax, ay, az);
            force_accumulate(dx1, dy1, dz1, r21, __fmul_rn(c1.w, mscale), ax, ay, az);
            if (COUNT && mine) nInter += 2 * nact;
        }
        if (nbody & 1) {
            const float4 c0 = lds_v4(brow + 16u * (unsigned)(nbody - 1));
            BH_DIST(c0, 0)
            force_accumulate(dx0, dy0, dz0, r20, __fmul_rn(c0.w, mscale), ax, ay, az);
            if (COUNT && mine) nInter += nact;
        }
#undef BH_CELL_VOTE
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            nInter += __shfl_xor_sync(kFull, nInter, o);
            nOpen += __shfl_xor_sync(kFull, nOpen, o);
        }
        if (lane == 0) {
            atomicAdd(&sc->interactions, nInter);
            atomicAdd(&sc->opens, nOpen);
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (t >= nact) break;
        const float4 a = make_float4(t ? ax.y : ax.x, t ? ay.y : ay.x, t ? az.y : az.x, 0.0f);
        for (int r = 0; r < dst.count; ++r) dst.buf[r][phase + k0 + t] = a;  // own buffer, then the peers' (NVLink stores)
    }
    __syncwarp();
    }
}
#undef BH_DIST

// ---- 6. finish: velocity correction + integrate + physical reordering --------------------------------------------
// calculateforce.cl:174-185 (vel += (a - a_old) * dt * 0.5 for step > 0; acc = a) and integrate.cl:27-43, for the
// body in slot perm[k], which then moves to slot k of the other buffers: after the step the bodies lie in this
// step's tree order, which is next step's insertion order (build), gather order (summarise) and walk order.
// APPLY = false: integrate only (the stage-by-stage API has applied the accelerations in calculate_force);
// PERMUTE = false: bodies stay where they are (bh_set_insertion_order(0), or no fresh sort).
// vpos/vvel (optional) = copyvertices.cl:14-17 fused: float4 {x,y,z,1} / {vx,vy,vz,1} in the host's numbering.
template <bool APPLY, bool PERMUTE>
__global__ void __launch_bounds__(256) finish_kernel(const float4 *bodyIn, const float4 *vaIn, float4 *bodyOut, float4 *vaOut,
                                                     const float4 *__restrict__ acc, unsigned phaseStride,
                                                     const int *__restrict__ perm, const Scalars *__restrict__ sc, int n, float dt,
                                                     float4 *__restrict__ vpos, float4 *__restrict__ vvel) {
    if (sc->error != 0) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int src = ((APPLY || PERMUTE) && perm) ? perm[k] : k;  // perm == nullptr: bodies already lie in tree order
    const int dstSlot = PERMUTE ? k : src;
    float4 p = bodyIn[src];
    float4 v = vaIn[2 * (size_t)src];
    float4 a = vaIn[2 * (size_t)src + 1];
    if (APPLY) {
        const int step = sc->step;
        const float4 an = acc[(size_t)(step & 1) * phaseStride + k];
        if (step > 0) {  // calculateforce.cl:174-179
            v.x = __fadd_rn(v.x, __fmul_rn(__fmul_rn(__fsub_rn(an.x, a.x), dt), 0.5f));
            v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(__fsub_rn(an.y, a.y), dt), 0.5f));
            v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(__fsub_rn(an.z, a.z), dt), 0.5f));
        }
        a = make_float4(an.x, an.y, an.z, 0.0f);  // :183-185
    }
    // integrate.cl:29-43
    const float dvx = __fmul_rn(__fmul_rn(a.x, dt), 0.5f);
    const float dvy = __fmul_rn(__fmul_rn(a.y, dt), 0.5f);
    const float dvz = __fmul_rn(__fmul_rn(a.z, dt), 0.5f);
    v.x = __fadd_rn(v.x, dvx); v.y = __fadd_rn(v.y, dvy); v.z = __fadd_rn(v.z, dvz);
    p.x = fmaf(v.x, dt, p.x); p.y = fmaf(v.y, dt, p.y); p.z = fmaf(v.z, dt, p.z);
    v.x = __fadd_rn(v.x, dvx); v.y = __fadd_rn(v.y, dvy); v.z = __fadd_rn(v.z, dvz);
    bodyOut[dstSlot] = p;
    vaOut[2 * (size_t)dstSlot] = v;  // v.w = origId travels with the body
    if (APPLY || PERMUTE) vaOut[2 * (size_t)dstSlot + 1] = a;
    if (vpos || vvel) {
        const int id = __float_as_int(v.w);
        if (vpos) vpos[id] = make_float4(p.x, p.y, p.z, 1.0f);
        if (vvel) vvel[id] = make_float4(v.x, v.y, v.z, 1.0f);
    }
}

// calculateforce.cl:174-185 alone (bh_calculate_force of the stage-by-stage API): in place
__global__ void __launch_bounds__(256) apply_acc_kernel(const float4 *__restrict__ acc, unsigned phaseStride,
                                                        const int *__restrict__ perm, float4 *__restrict__ velacc,
                                                        const Scalars *__restrict__ sc, int n, float dt) {
    if (sc->error != 0) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int slot = perm ? perm[k] : k;
    const int step = sc->step;
    const float4 a = acc[(size_t)(step & 1) * phaseStride + k];
    if (step > 0) {
        float4 v = velacc[2 * (size_t)slot];
        const float4 a0 = velacc[2 * (size_t)slot + 1];
        v.x = __fadd_rn(v.x, __fmul_rn(__fmul_rn(__fsub_rn(a.x, a0.x), dt), 0.5f));
        v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(__fsub_rn(a.y, a0.y), dt), 0.5f));
        v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(__fsub_rn(a.z, a0.z), dt), 0.5f));
        velacc[2 * (size_t)slot] = v;
    }
    velacc[2 * (size_t)slot + 1] = make_float4(a.x, a.y, a.z, 0.0f);
}

// ---- multi-GPU: device-side barrier over peer memory ---------------------------------------------------------------
// Every rank owns `flags[kMaxPeers]` (unsigned long long) in its peer-mapped allocation.  barrier_kernel (one
// warp) bumps the rank's own sequence number, release-stores it at system scope into slot `rank` of every peer's
// flags, and waits until every peer's number has arrived in its own flags.  Stream order puts the walk's peer
// stores before the barrier's release and the barrier's acquire before finish_kernel's loads.
struct PeerFlags {
    unsigned long long *flags[kMaxPeers];  // flags[r] = rank r's flag array (own = local pointer)
    int count, rank;
};

__global__ void barrier_kernel(const PeerFlags pf, unsigned long long *seq, Scalars *sc, long long timeoutCycles) {
    const int r = threadIdx.x;
    unsigned long long s = *seq + 1;
    __syncwarp();
    if (r == 0) *seq = s;
    if (r >= pf.count || r == pf.rank) return;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pf.flags[r] + pf.rank), "l"(s) : "memory");
    const unsigned long long *mine = pf.flags[pf.rank] + r;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= s) break;
        if (clock64() - t0 > timeoutCycles) {  // a peer died or fell out of step: report instead of hanging
            atomicCAS(&sc->error, 0, 3);
            break;
        }
        __nanosleep(64);
    }
}

// ---- host-boundary helpers: pack uploads, export logical buffers -------------------
// positions + masses (what the tree stages and the walk read), and velocities + the host's numbering (what only the
// finish pass reads): two kernels so that an asynchronous upload can deliver the second half while the step already runs
__global__ void pack_pos_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                                const float *__restrict__ mass, float4 *__restrict__ body4, const int *__restrict__ slotOf,
                                int *__restrict__ perm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = slotOf ? slotOf[i] : i;
    if ((unsigned)slot < (unsigned)n) body4[slot] = make_float4(x[i], y[i], z[i], mass[i]);
    perm[i] = 0;
}

__global__ void pack_vel_kernel(const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                                float4 *__restrict__ velacc, const int *__restrict__ slotOf, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = slotOf ? slotOf[i] : i;
    if ((unsigned)slot >= (unsigned)n) return;
    velacc[2 * (size_t)slot] = make_float4(vx[i], vy[i], vz[i], __int_as_float(i));  // one whole 32-byte sector per body
    velacc[2 * (size_t)slot + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// Where an upload puts body i: the slot the body with the host's number i occupies in the state being replaced
// (tree order of the last step).  A host that sends its bodies every step -- the same bodies, a little further on --
// then hands the tree stages bodies that are already nearly in tree order, as in a run that never leaves the
// device.  Any placement is correct (the host's numbering travels with the body, see origId); this one is fast.
__global__ void slot_of_kernel(const float4 *__restrict__ velacc, int *__restrict__ slotOf, int n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int id = __float_as_int(velacc[2 * (size_t)s].w);
    if ((unsigned)id < (unsigned)n) slotOf[id] = s;
}

// A reset (upload, generated universe) returns the tree buffers to the reference's initial state, all zeros
// (GPUBH:155-179).  Only the cells builds have allocated since the last reset can differ from that.
__global__ void clear_tree_kernel(int *__restrict__ child, int *__restrict__ start, int *__restrict__ count,
                                  float4 *__restrict__ cell4, const Scalars *sc, int n, int m) {
    if (sc->lowWater == kNothingDirty) return;  // no stage has run since the last reset
    int lo = min(sc->lowWater, sc->bottom);
    if (sc->error != 0 || lo < n) lo = n;  // a failed build: take no chances
    const int4 zero = make_int4(0, 0, 0, 0);
    for (long long c = lo - n + blockIdx.x * (long long)blockDim.x + threadIdx.x; c <= m - n; c += gridDim.x * (long long)blockDim.x) {
        reinterpret_cast<int4 *>(child)[2 * c] = zero;
        reinterpret_cast<int4 *>(child)[2 * c + 1] = zero;
        start[c] = 0;
        count[c] = 0;
        cell4[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// Logical float buffers (GPUBH:198-205): bodies 0..n-1 in the host's numbering, cells n..m; `which`: 0-2 pos,
// 3-5 vel, 6-8 acc, 9 mass.  vel/acc are zero beyond the bodies (the reference never writes there).
__global__ void export_float_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ velacc,
                                    const float4 *__restrict__ cell4, int which, float *__restrict__ out, int n, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len && i >= n) return;
    if (i < n) {  // slot i: scatter to the host's index
        const float4 v0 = velacc[2 * (size_t)i];
        const int id = __float_as_int(v0.w);
        if (id < len) {
            float r;
            if (which < 3 || which == 9) {
                const float4 p = body4[i];
                r = which == 0 ? p.x : which == 1 ? p.y : which == 2 ? p.z : p.w;
            } else if (which < 6) {
                r = which == 3 ? v0.x : which == 4 ? v0.y : v0.z;
            } else {
                const float4 a = velacc[2 * (size_t)i + 1];
                r = which == 6 ? a.x : which == 7 ? a.y : a.z;
            }
            out[id] = r;
        }
    }
    if (i >= n && i < len) {
        float r = 0.0f;
        if (which < 3 || which == 9) {
            const float4 c = cell4[i - n];
            r = which == 0 ? c.x : which == 1 ? c.y : which == 2 ? c.z : c.w;
        }
        out[i] = r;
    }
}

// logical int arrays that exist only for cells: zeros for the first n entries; mask = bits to keep
__global__ void export_shifted_kernel(const int *__restrict__ src, long long skip, int mask, int *__restrict__ out, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len) return;
    int v = 0;
    if (i >= skip) {
        v = src[i - skip];
        if (v >= 0) v &= mask;
    }
    out[i] = v;
}

// child[8(M+1)]: zeros for the body rows; body slots translated to the host's numbering through the origIds of the
// buffers the tree was built from.  Rows of cells no build has allocated (below `bottom`, or all of them after a
// reset) hold the initial zeros, which are not body slots.
__global__ void export_child_kernel(const int *__restrict__ child, const float4 *__restrict__ velaccTree, const Scalars *sc, int n,
                                    int *__restrict__ out, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len) return;
    int v = 0;
    if (i >= 8ll * n) {
        v = child[i - 8ll * n];
        const bool allocated = sc->lowWater != kNothingDirty && (i >> 3) >= sc->bottom;
        if (allocated && v >= 0 && v < n) v = __float_as_int(velaccTree[2 * (size_t)v].w);
    }
    out[i] = v;
}

// sorted[M+1]: pending = the sort's permutation has not been applied yet: origId[perm[k]]; else the bodies lie in
// tree order: origId[k].  Zeros beyond the bodies.
__global__ void export_sorted_kernel(const int *__restrict__ perm, const float4 *__restrict__ velacc, int pending, int n,
                                     int *__restrict__ out, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len) return;
    int v = 0;
    if (i < n) v = __float_as_int(velacc[2 * (size_t)(pending ? perm[i] : (int)i)].w);
    out[i] = v;
}

// copyvertices.cl:14-17, host numbering
__global__ void copy_vertices_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ velacc,
                                     float4 *__restrict__ pos, float4 *__restrict__ vel, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = velacc[2 * (size_t)i];
    const int id = __float_as_int(v.w);
    if (pos) {
        const float4 p = body4[i];
        pos[id] = make_float4(p.x, p.y, p.z, 1.0f);
    }
    if (vel) vel[id] = make_float4(v.x, v.y, v.z, 1.0f);
}

// The same by gather: vertex i comes from the slot of body i.  Whole 16-byte vertices are written in order; the
// scattered side is the reads (a 32-byte sector per 16 bytes used), which cost a third of what scattered 16-byte
// writes do (each of those makes L2 fetch the rest of its sector from DRAM first).
__global__ void copy_vertices_gather_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ velacc,
                                            const int *__restrict__ slotOf, float4 *__restrict__ pos, float4 *__restrict__ vel, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = slotOf[i];
    if ((unsigned)slot >= (unsigned)n) return;
    if (pos) {
        const float4 p = body4[slot];
        pos[i] = make_float4(p.x, p.y, p.z, 1.0f);
    }
    if (vel) {
        const float4 v = velacc[2 * (size_t)slot];
        vel[i] = make_float4(v.x, v.y, v.z, 1.0f);
    }
}

// SoA state dump in the host's numbering (bh_write_universe_file): out = 7 arrays of n floats
__global__ void export_universe_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ velacc, float *__restrict__ out,
                                       int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = body4[i], v = velacc[2 * (size_t)i];
    const size_t id = (size_t)__float_as_int(v.w), N = (size_t)n;
    out[id] = p.x; out[N + id] = p.y; out[2 * N + id] = p.z;
    out[3 * N + id] = v.x; out[4 * N + id] = v.y; out[5 * N + id] = v.z;
    out[6 * N + id] = p.w;
}

// ---- seeded universe generators on the device (ch.fhnw.woipv.nbody.simulation.universe.*) -----------------------
// The reference's generators draw from the unseeded Math.random(); these draw from Philox4x32-10, one
// subsequence per body, so a universe is a pure function of (kind, seed, n) and needs no host memory.
// kind 0: RandomCubicUniverseGenerator.java:13-17   (U-0.5)*range per axis, v = 0, m = 1/n          (p0 = range)
// kind 1: PlummerUniverseGenerator.java:8-41        Plummer sphere, mass-fraction cut 0.999, m = 1/n
// kind 2: RotatingDiskGalaxyGenerator.java:17-43    disk of radius p0, velocity multiplier p1, body 0 = centre mass p2
__device__ __forceinline__ double philox_uniform(curandStatePhilox4_32_10_t *st) { return 1.0 - curand_uniform_double(st); }  // [0,1)

__global__ void __launch_bounds__(256) generate_kernel(float4 *__restrict__ body4, float4 *__restrict__ velacc,
                                                       int *__restrict__ perm, int n, int kind, unsigned long long seed,
                                                       float p0, float p1, float p2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, 0, &st);
    float x = 0.f, y = 0.f, z = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, mass = 1.0f / (float)n;
    if (kind == 0) {
        x = (float)((philox_uniform(&st) - 0.5) * p0);
        y = (float)((philox_uniform(&st) - 0.5) * p0);
        z = (float)((philox_uniform(&st) - 0.5) * p0);
    } else if (kind == 1) {
        const double kPi = 3.14159265358979323846;
        const double rsc = (3 * kPi) / 16, vsc = sqrt(1.0 / rsc);
        mass = (float)(1.0 / n);
        const double r = 1.0 / sqrt(pow(philox_uniform(&st) * 0.999, -2.0 / 3.0) - 1);
        double a, b, c, sq;
        do {
            a = philox_uniform(&st) * 2.0 - 1.0; b = philox_uniform(&st) * 2.0 - 1.0; c = philox_uniform(&st) * 2.0 - 1.0;
            sq = a * a + b * b + c * c;
        } while (sq > 1.0 || sq == 0.0);
        double scale = rsc * r / sqrt(sq);
        x = (float)(a * scale); y = (float)(b * scale); z = (float)(c * scale);
        do {
            a = philox_uniform(&st); b = philox_uniform(&st) * 0.1;
        } while (b > a * a * pow(1 - a * a, 3.5));
        const double v = a * sqrt(2.0 / sqrt(1 + r * r));
        do {
            a = philox_uniform(&st) * 2.0 - 1.0; b = philox_uniform(&st) * 2.0 - 1.0; c = philox_uniform(&st) * 2.0 - 1.0;
            sq = a * a + b * b + c * c;
        } while (sq > 1.0 || sq == 0.0);
        scale = vsc * v / sqrt(sq);
        vx = (float)(a * scale); vy = (float)(b * scale); vz = (float)(c * scale);
    } else {
        if (i == 0) {
            mass = p2;  // bodiesMass[0] = centerMass, at the origin
        } else {
            const float r = (float)(philox_uniform(&st) * p0) + 0.05f;
            const double alpha = philox_uniform(&st) * 2 * 3.14159265358979323846;
            x = (float)(cos(alpha) * r); y = (float)(sin(alpha) * r);
            z = (float)((philox_uniform(&st) - 0.5) / 8);
            const float v0 = (float)sqrt((double)((p2 + mass) / (r * r * r))) * p1;
            vx = y * v0; vy = -x * v0;
        }
    }
    body4[i] = make_float4(x, y, z, mass);
    velacc[2 * (size_t)i] = make_float4(vx, vy, vz, __int_as_float(i));
    velacc[2 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    perm[i] = 0;
}

// ---- diagnostics (GPUBH:305-365 printEnergy / printImpulse, on the device) ---------------------------
// out[0] = sum 1/2 m v^2, out[1..3] = sum m v, out[4] = sum m; double accumulation.
__global__ void __launch_bounds__(256) kinetic_kernel(const float4 *__restrict__ body4, const float4 *__restrict__ velacc,
                                                      double *__restrict__ out, int n) {
    double e = 0.0, px = 0.0, py = 0.0, pz = 0.0, ms = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float m = body4[i].w;
        const float4 v = velacc[2 * (size_t)i];
        e += 0.5 * m * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
        px += (double)m * v.x; py += (double)m * v.y; pz += (double)m * v.z; ms += m;
    }
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o); px += __shfl_xor_sync(0xffffffffu, px, o);
        py += __shfl_xor_sync(0xffffffffu, py, o); pz += __shfl_xor_sync(0xffffffffu, pz, o);
        ms += __shfl_xor_sync(0xffffffffu, ms, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 0, e); atomicAdd(out + 1, px); atomicAdd(out + 2, py); atomicAdd(out + 3, pz); atomicAdd(out + 4, ms);
    }
}

// out[5] += -1/2 sum_i m_i sum_{j != i} m_j / sqrt(r_ij^2 + eps): the softened potential of the parity tests
// (not the reference's unsoftened, doubled printEnergy term).  Direct sum, j-tiles staged through shared memory,
// fp32 pair terms, per-tile fp32 partial sums folded into double.
constexpr int kPotTile = 256;
__global__ void __launch_bounds__(kPotTile) potential_kernel(const float4 *__restrict__ body4, double *__restrict__ out, int n,
                                                             float eps) {
    __shared__ float4 tile[kPotTile];
    const int i = blockIdx.x * kPotTile + threadIdx.x;
    const float4 pi = body4[min(i, n - 1)];
    double phi = 0.0;
    for (int j0 = 0; j0 < n; j0 += kPotTile) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < n ? body4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        float part = 0.0f;
#pragma unroll 8
        for (int t = 0; t < kPotTile; ++t) {
            const float4 pj = tile[t];
            const float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) + eps;
            const float w = (j0 + t == i) ? 0.0f : pj.w;  // no self term; padded slots have mass 0
            part = fmaf(w, rsqrtf(r2), part);
        }
        phi += part;
        __syncthreads();
    }
    double e = i < n ? -0.5 * (double)pi.w * phi : 0.0;
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + 5, e);
}

// Tree potential (bh_diagnostics mode 2): the walk's POT variant leaves S_k = sum_j m_j / sqrt(r_kj^2 + eps) over the nodes
// it used -- the body itself included, at r = 0 -- in acc[k].w; out[5] += -1/2 sum_k m_k (S_k - m_k / sqrt(eps)).
__global__ void __launch_bounds__(256) tree_potential_kernel(const float4 *__restrict__ body4, const int *__restrict__ perm,
                                                             const float4 *__restrict__ acc, unsigned phaseStride,
                                                             const Scalars *__restrict__ sc, double *__restrict__ out, int n, float eps) {
    const float4 *a = acc + (size_t)(sc->step & 1) * phaseStride;
    const float selfInv = rsqrtf(eps);
    double e = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const float m = body4[perm ? perm[k] : k].w;
        e += -0.5 * (double)m * ((double)a[k].w - (double)m * (double)selfInv);
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + 5, e);
}

__global__ void adjust_step_kernel(Scalars *sc, int delta) { sc->step += delta; }

// Small device-side resets as kernels, not cudaMemsetAsync / cudaMemcpyAsync: those may be queued on a copy engine,
// where they wait behind a bulk host transfer of the asynchronous boundary (measured: ~0.9 ms per step while a
// 320 MB read-back was in flight).
__global__ void reset_ticket_kernel(Scalars *sc) {
    sc->walkTicket = 0;
    sc->walkSpills = 0;
}
__global__ void reset_counters_kernel(Scalars *sc) {
    sc->interactions = 0;
    sc->opens = 0;
}
__global__ void init_scalars_kernel(Scalars *sc) {  // GPUBH:155-179: everything zero except step = -1, maxDepth = 1
    Scalars init = {};
    init.step = -1;
    init.maxDepth = 1;
    init.lowWater = kNothingDirty;
    *sc = init;
}

// ---- measurement utility: FP32 FMA peak of the device (roofline denominator of the force kernel) ----
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456f) out[0] = r;  // never true; keeps the chains alive
}

// The same with the instruction the walk is made of: FFMA2 (packed fp32x2) with three distinct register-pair operands per
// instruction (no constants, no immediate, no operand reuse): what the register file lets the fp32 pipe sustain for real code.
__global__ void __launch_bounds__(256) fp32x2_peak_kernel(float *out, int iters, float a, float b) {
    float2 x[8], y[8], z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        x[u] = make_float2(threadIdx.x + u, threadIdx.x - u);
        y[u] = make_float2(a + 1e-7f * u, a - 1e-7f * u);
        z[u] = make_float2(b * (u + 1), b * (u + 2));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ffma2_rn(x[u], y[(u + r) & 7], z[(u + 3 * r) & 7]);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) r += x[u].x + x[u].y;
    if (r == 123.456f) out[0] = r;  // never true; keeps the chains alive
}

}  // namespace bh
