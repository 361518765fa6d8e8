// bh_kernels.cuh -- hand-written sm_100a kernels of the Barnes-Hut step.
//
// One kernel per stage of the reference's step (GPUBarnesHutNBodySimulation.java:258-263):
//   bbox_kernel       <- kernels/nbody/boundingbox.cl
//   build_kernel      <- kernels/nbody/buildtree.cl
//   summarize_kernel  <- kernels/nbody/summarizetree.cl
//   sort_kernel       <- kernels/nbody/sort.cl
//   force2_kernel     <- kernels/nbody/calculateforce.cl
//   integrate_kernel  <- kernels/nbody/integrate.cl
// They reproduce the reference's *results* (see DESIGN.md for the parity classes),
// not its code: the data layout, work decomposition and synchronisation are
// designed for B200.
//
// HBM layout (N bodies, M = number of nodes, NC = M - N + 1 cell slots):
//   node4  float4[M+1]   {x, y, z, mass}; bodies 0..N-1, cells N..M, root = M.
//                        During build a cell holds its geometric centre and
//                        mass = -1; summarise overwrites it with {COM, mass}.
//   velacc float4[2N]    {vx,vy,vz,0},{ax,ay,az,0} per body: one 32-byte sector.
//   child  int[8*NC]     child[(cell-N)*8 + k]; -1 empty, -2 locked, <N body, >=N cell
//   octet  float4[8*NC]  the force walk's record of a cell: copies of its children's node4
//   oidx   int[8*NC]     records, child cells first, then child bodies (128-byte line), the
//   meta   int[NC]       child cells' indices minus N, and #cells | #bodies << 4.  Written by summarise.
//   start, count int[NC] `start` and `bodyCount` of the reference; count doubles
//                        as the "summarised" flag (-1 = not yet).
//   sorted int[N]        bodies in tree (DFS) order.
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

struct Scalars {
    int step;         // init -1 (GPUBH:165)
    int blockCount;   // last-block-done ticket of bbox_kernel
    float radius;
    int maxDepth;     // init 1, running max (buildtree.cl:199)
    int bottom;
    int error;
    int pad0, pad1;
    unsigned long long interactions;
    unsigned long long opens;
};

// ---- memory-model helpers -------------------------------------------------
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

__global__ void __launch_bounds__(kBboxThreads) bbox_kernel(float4 *__restrict__ node4, int *__restrict__ child,
                                                            int *__restrict__ start, int *__restrict__ count,
                                                            int *__restrict__ arrived, float *__restrict__ partials,
                                                            Scalars *__restrict__ sc, int n, int m) {
    __shared__ float red[6][kBboxThreads / 32];
    __shared__ bool isLast;
    const float4 seed = node4[0];  // boundingbox.cl:44-58: every lane starts from body 0
    float mnx = seed.x, mny = seed.y, mnz = seed.z, mxx = seed.x, mxy = seed.y, mxz = seed.z;
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
        const float4 p0 = node4[i], p1 = node4[i + stride], p2 = node4[i + 2 * stride], p3 = node4[i + 3 * stride];
        mnx = fminf(fminf(mnx, p0.x), fminf(fminf(p1.x, p2.x), p3.x)); mxx = fmaxf(fmaxf(mxx, p0.x), fmaxf(fmaxf(p1.x, p2.x), p3.x));
        mny = fminf(fminf(mny, p0.y), fminf(fminf(p1.y, p2.y), p3.y)); mxy = fmaxf(fmaxf(mxy, p0.y), fmaxf(fmaxf(p1.y, p2.y), p3.y));
        mnz = fminf(fminf(mnz, p0.z), fminf(fminf(p1.z, p2.z), p3.z)); mxz = fmaxf(fmaxf(mxz, p0.z), fmaxf(fmaxf(p1.z, p2.z), p3.z));
    }
    for (; i < n; i += stride) {
        const float4 p = node4[i];
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
        sc->bottom = m;
        node4[m] = make_float4(rx, ry, rz, -1.0f);
        start[m - n] = 0;
        count[m - n] = -1;
        arrived[m - n] = 0;
        sc->step = sc->step + 1;
    }
}

// ---- 2. tree build ------------------------------------------------------------
// buildtree.cl: concurrent insertion; a child slot is locked by CAS to -2 while a
// leaf is split, the finished sub-tree is published after a device fence.  The
// tree *shape* is a function of the positions and the root box only; cell
// numbers depend on the allocation race (as in the reference).  Bodies are
// visited through `order` (previous step's sorted[] = spatial order) when given.
//
// B200 design, all of it about latency (the kernel issues < 20 % of its slots):
//  * Every lane owns a contiguous run of the insertion order, so the body it
//    inserts next is its spatial neighbour.  The lane remembers the path of its
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
constexpr int kBuildThreads = 256;
constexpr int kPathCap = 24;  // remembered levels per lane (24 KB of shared memory per CTA); deeper levels are loaded

__global__ void __launch_bounds__(kBuildThreads) build_kernel(float4 *__restrict__ node4, int *child,
                                                              int *__restrict__ start, int *__restrict__ count,
                                                              int *__restrict__ parent, int *__restrict__ arrived,
                                                              const int *__restrict__ order, Scalars *sc, int n, int m) {
    constexpr unsigned kFull = 0xffffffffu;
    const float radius = sc->radius;
    const float4 root = node4[m];
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
                body = order ? order[i] : i;
                p = node4[body];
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
                    q = node4[ch];
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
                        node4[cell] = make_float4(cx, cy, cz, -1.0f);
                        start[cell - n] = -1;
                        count[cell - n] = -1;
                        parent[cell - n] = cur;  // for the counter-driven summarise
                        const int qPath = octant(cx, cy, cz, q.x, q.y, q.z);
                        const int pPath = octant(cx, cy, cz, p.x, p.y, p.z);
                        arrived[cell - n] = (qPath == pPath) ? (1 << 16) : 0;  // (#child cells << 16) | reports received
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
                    atomicAdd(arrived + (node0 - n), 1 << 16);  // the leaf's cell gains a child cell
                    __threadfence();          // :173 publish the sub-tree ...
                    st_relaxed(slot, patch);  // :180 ... by replacing the lock
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
// own thread (`arrived`: build keeps the number of child cells in the high half,
// reports count up in the low half; a small array that stays in L2); whoever
// reports last summarises the cell, reports to the parent and climbs on.  Children are summed in octant order,
// which makes the result independent of timing and bit-identical to the oracle.
// The summarising thread also writes the force walk's record of the cell.
constexpr int kSummThreads = 256;

__global__ void __launch_bounds__(kSummThreads, 4) summarize_kernel(float4 *__restrict__ node4, int *__restrict__ child,
                                                                 float4 *__restrict__ octet, int *__restrict__ oidx,
                                                                 int *__restrict__ meta, int *__restrict__ count,
                                                                 const int *__restrict__ parent, int *arrived, Scalars *sc,
                                                                 int n, int m) {
    if (sc->error != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->bottom = m;  // buildtree.cl:117
        return;
    }
    const int bottom = sc->bottom;
    const int stride = gridDim.x * blockDim.x;
    for (int first = bottom + blockIdx.x * blockDim.x + threadIdx.x; first <= m; first += stride) {
        int cell = first;
        // A cell is summarised by whoever reports last among its child cells and its own thread.  `arrived` holds
        // the number of child cells (maintained by build) in its high half and the reports in its low half, so one
        // atomic on a small, L2-resident array both reports and tells whether this was the last report.
        int old = atomicAdd(arrived + (cell - n), 1);
        if ((old & 0xffff) != (old >> 16)) continue;
        __threadfence();
        for (;;) {
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
                    c[k] = __ldcg(node4 + out[k]);
                    cnt[k] = __ldcg(count + (out[k] - n));
                } else if (out[k] >= 0) {
                    c[k] = node4[out[k]];
                    cnt[k] = 1;
                }
            }
            float cm = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
            int bodies = 0;  // summarizetree.cl:98-105,118
            float4 *orow = octet + (size_t)(cell - n) * 8;
            int *irow = oidx + (size_t)(cell - n) * 8;
            int cpos = 0, bpos = ncell;  // the force walk's record: child cells first, then child bodies
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (out[k] < 0) break;
                bodies += cnt[k];
                cm = __fadd_rn(cm, c[k].w);  // summarizetree.cl:107-110, octant order
                cx = fmaf(c[k].x, c[k].w, cx);
                cy = fmaf(c[k].y, c[k].w, cy);
                cz = fmaf(c[k].z, c[k].w, cz);
                if (out[k] >= n) {
                    irow[cpos] = out[k] - n;
                    orow[cpos++] = c[k];
                } else {
                    orow[bpos++] = c[k];
                }
            }
            reinterpret_cast<int4 *>(row)[0] = make_int4(out[0], out[1], out[2], out[3]);
            reinterpret_cast<int4 *>(row)[1] = make_int4(out[4], out[5], out[6], out[7]);
            meta[cell - n] = ncell | ((used - ncell) << 4);  // force walk: #child cells, #child bodies
            const float inv = __frcp_rn(cm);  // summarizetree.cl:161: 1.0f / cellMass, correctly rounded
            __stcg(node4 + cell, make_float4(__fmul_rn(cx, inv), __fmul_rn(cy, inv), __fmul_rn(cz, inv), cm));
            __stcg(count + (cell - n), bodies);  // summarizetree.cl:160
            if (cell == m) break;                // the root
            // report to the parent; the child that reports last continues with it
            const int par = parent[cell - n];
            __threadfence();  // summarizetree.cl:170: this cell's record before the report
            old = atomicAdd(arrived + (par - n), 1);
            if ((old & 0xffff) != (old >> 16)) break;  // not the last report: somebody else continues
            __threadfence();
            cell = par;
        }
    }
}

// ---- 4. sort ----------------------------------------------------------------------
// sort.cl: top-down propagation of `start`, bodies written in DFS order.  One
// thread per cell, descending index (parents first), waits for start >= 0.
constexpr int kSortThreads = 256;

__global__ void __launch_bounds__(kSortThreads) sort_kernel(const int *__restrict__ child, const int *__restrict__ count,
                                                            int *start, int *__restrict__ sorted, Scalars *sc, int n, int m) {
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
                s += count[ch[k] - n];
            } else {               // sort.cl:59-65
                sorted[s++] = ch[k];
            }
        }
    }
}

constexpr int kStackCap = 7 * kMaxDepth + 8;  // a popped cell pushes at most 8 children, 7 stay while the 8th is walked

// ---- 5. force -------------------------------------------------------------------
// calculateforce.cl: a vote group of VOTE consecutive sorted bodies walks the tree
// together; a cell is used as a point mass only if *all* bodies of the group
// are far enough (work_group_all, :145), bodies are always used.  The set of
// (group, node) interactions is exactly the reference's; the order in which a
// body sums them differs (all children of a popped cell are consumed before its
// opened children are descended), which moves the fp32 sum by rounding only.
// The kernel is built around Blackwell's packed fp32 pipe
// (FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issue slot).  The walk is
// issue-bound, so every lane carries TWO consecutive sorted bodies and a warp
// carries 64 bodies = 64/VOTE vote groups (lanes 8g..8g+7 are group g for
// VOTE = 16).  A stack entry is {cell - N, one bit per group that still needs the
// cell | depth << 1}; a group that accepted a cell is simply absent from the
// mask of its children.  Per pop one LDG.128 brings the cell's walk record (8
// children + 8 indices, written by summarise) into a per-warp shared-memory row
// -- per-child global loads miss L1 on every second child (32-byte sectors) and
// each miss costs an L2 round trip -- and the children are consumed with
// broadcast LDS.128: child cells first (7 packed fp32 ops, one ballot; a second
// ballot and the push only if some body is too near), then child bodies.  (An L1
// prefetch of pushed cells was measured and dropped: 2 % slower than none.)
constexpr int kForce2Threads = 128;
constexpr int kForce2Bodies = 2 * kForce2Threads;  // per CTA

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
__device__ __forceinline__ float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }

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

// Destinations of a slice's accelerations: the rank's own sorted-order buffer and, in a multi-GPU run with
// peer memory, the same buffer of every other rank mapped over NVLink (CUDA IPC): the walk's epilogue stores
// straight into all of them, so the all-gather is fused into the force kernel and overlaps the walk.
constexpr int kMaxPeers = 16;
struct PeerBuffers {
    float4 *buf[kMaxPeers];
    int count;
};

template <int VOTE, bool SLICE, bool COUNT>
__global__ void __launch_bounds__(kForce2Threads) force2_kernel(const float4 *__restrict__ node4, const float4 *__restrict__ octet,
                                                                 const int *__restrict__ oidx, const int *__restrict__ meta,
                                                                 const int *__restrict__ sorted, float4 *__restrict__ velacc,
                                                                 const PeerBuffers dst, Scalars *sc, int n, int m,
                                                                 int first, int cnt, float thetaMacro, float eps, float dt) {
    __shared__ float dq[kMaxDepth];
    __shared__ int2 stack[kForce2Threads / 32][kStackCap];  // {cell - N, group bits | depth << 1}
    __shared__ float4 stage[kForce2Threads / 32][10];       // the popped cell's walk record: 8 children, 8 indices
    if (sc->error != 0) return;
    const int maxDepth = sc->maxDepth;
    if (maxDepth > kMaxDepth) {  // calculateforce.cl:69-73
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->error = 1;
        return;
    }
    if (threadIdx.x == 0) {  // calculateforce.cl:52-67
        const float radius = sc->radius;
        float v = __fmul_rn(radius, radius);
        if (thetaMacro > 0.0f) v = __fdiv_rn(v, thetaMacro);
        for (int i = 0; i < maxDepth; ++i) {
            dq[i] = __fadd_rn(v, eps);
            v = __fmul_rn(0.25f, v);
        }
    }
    __syncthreads();
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kLanesPerGroup = VOTE / 2;  // two bodies per lane
    constexpr unsigned kGroupLanes = (kLanesPerGroup == 32) ? kFull : ((1u << kLanesPerGroup) - 1u);
    constexpr unsigned kSpread = (kLanesPerGroup == 8) ? 0x01010101u : (kLanesPerGroup == 16) ? 0x00010001u : 1u;  // first lane of every group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int end = first + cnt;
    const int base = first + (blockIdx.x * (kForce2Threads / 32) + warp) * 64;
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
    const int b0 = sorted[s0], b1 = sorted[s1];
    const float4 p0 = node4[b0], p1 = node4[b1];
    const float2 npx = make_float2(-p0.x, -p1.x), npy = make_float2(-p0.y, -p1.y), npz = make_float2(-p0.z, -p1.z);
    const float2 eps2 = make_float2(eps, eps);
    float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
    unsigned long long nInter = 0, nOpen = 0;
    // Shared memory is addressed through 32-bit shared-window addresses kept in registers (ptxas otherwise
    // rebuilds the window base from SR_CgaCtaId on every pop).
    const unsigned stkBase = (unsigned)__cvta_generic_to_shared(stack[warp]);
    unsigned rowBase = (unsigned)__cvta_generic_to_shared(stage[warp]);
    unsigned dqBase = (unsigned)__cvta_generic_to_shared(dq);
    unsigned rowLane = rowBase + 16u * (unsigned)lane;
    // lanes 0-7 fetch the 8 child records, lanes 8-9 the 8 child indices
    const char *laneBase = lane < 8 ? reinterpret_cast<const char *>(octet + lane)
                                    : reinterpret_cast<const char *>(oidx) + 16 * ((lane - 8) & 1);
    int laneStride = lane < 8 ? 128 : 32;
    asm volatile("" : "+r"(rowBase), "+r"(dqBase), "+r"(rowLane), "+r"(laneStride));  // keep them in registers
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
        if (lane < 10) sts_v4(rowLane, __ldg(reinterpret_cast<const float4 *>(lane_address(laneBase, rel, laneStride))));
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
#define BH_DIST(c, S)                                                                                                  \
    const float2 dx##S = __fadd2_rn(make_float2((c).x, (c).x), npx); /* c - p, exactly */                              \
    const float2 dy##S = __fadd2_rn(make_float2((c).y, (c).y), npy);                                                   \
    const float2 dz##S = __fadd2_rn(make_float2((c).z, (c).z), npz);                                                   \
    const float2 r2##S = __fadd2_rn(__ffma2_rn(dz##S, dz##S, __ffma2_rn(dy##S, dy##S, __fmul2_rn(dx##S, dx##S))), eps2); /* :138-143 */
        // one child cell whose vote was not unanimous (or that was not tested as part of a pair)
#define BH_CELL_VOTE(c, S, far, j)                                                                                     \
    if (__all_sync(kFull, far)) { /* far enough for every body of the warp: every group that is here uses it */      \
        force_accumulate(dx##S, dy##S, dz##S, r2##S, __fmul_rn((c).w, mscale), ax, ay, az);                            \
        if (COUNT && mine) nInter += nact;                                                                             \
    } else {                                                                                                           \
        const unsigned near = __ballot_sync(kFull, !(far));                                                            \
        const unsigned open = __ballot_sync(kFull, (near & gmMine) != 0u);                                             \
        if (open) { /* :154-163 */                                                                                     \
            const int ch = lds_s32(rowBase + 128u + 4u * (j));                                                         \
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
        // Child cells come first and are taken two at a time: two independent distance chains per lane hide the
        // fp32 latency (dependency waits were the top stall), and the common case -- both cells far from every
        // body of the warp (the group vote of :145 is then unanimous in all groups) -- costs one VOTE for two.
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
#undef BH_DIST
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
    const bool corr = sc->step > 0;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (t >= nact) break;
        const int body = t ? b1 : b0;
        const float fx = t ? ax.y : ax.x, fy = t ? ay.y : ay.x, fz = t ? az.y : az.x;
        if (SLICE) {
            const float4 a = make_float4(fx, fy, fz, 0.0f);
            for (int r = 0; r < dst.count; ++r) dst.buf[r][k0 + t] = a;  // own buffer, then the peers' (NVLink stores)
        } else {
            if (corr) {  // calculateforce.cl:174-179
                float4 v = velacc[2 * (size_t)body];
                const float4 a0 = velacc[2 * (size_t)body + 1];
                v.x = __fadd_rn(v.x, __fmul_rn(__fmul_rn(__fsub_rn(fx, a0.x), dt), 0.5f));
                v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(__fsub_rn(fy, a0.y), dt), 0.5f));
                v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(__fsub_rn(fz, a0.z), dt), 0.5f));
                velacc[2 * (size_t)body] = v;
            }
            velacc[2 * (size_t)body + 1] = make_float4(fx, fy, fz, 0.0f);  // :183-185
        }
    }
}

// Multi-GPU: velocity correction + acc store from the all-gathered sorted-order
// accelerations (calculateforce.cl:174-185 for every body).
__global__ void __launch_bounds__(256) apply_acc_kernel(const float4 *__restrict__ accSorted, const int *__restrict__ sorted,
                                                        float4 *__restrict__ velacc, const Scalars *__restrict__ sc, int n,
                                                        float dt) {
    if (sc->error != 0) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int body = sorted[k];
    const float4 a = accSorted[k];
    if (sc->step > 0) {
        float4 v = velacc[2 * (size_t)body];
        const float4 a0 = velacc[2 * (size_t)body + 1];
        v.x = __fadd_rn(v.x, __fmul_rn(__fmul_rn(__fsub_rn(a.x, a0.x), dt), 0.5f));
        v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(__fsub_rn(a.y, a0.y), dt), 0.5f));
        v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(__fsub_rn(a.z, a0.z), dt), 0.5f));
        velacc[2 * (size_t)body] = v;
    }
    velacc[2 * (size_t)body + 1] = make_float4(a.x, a.y, a.z, 0.0f);
}

// ---- 6. integrate ---------------------------------------------------------------
// integrate.cl:27-43
__global__ void __launch_bounds__(256) integrate_kernel(float4 *__restrict__ node4, float4 *__restrict__ velacc,
                                                        const Scalars *__restrict__ sc, int n, float dt) {
    if (sc->error != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = node4[i];
    float4 v = velacc[2 * (size_t)i];
    const float4 a = velacc[2 * (size_t)i + 1];
    const float dvx = __fmul_rn(__fmul_rn(a.x, dt), 0.5f);
    const float dvy = __fmul_rn(__fmul_rn(a.y, dt), 0.5f);
    const float dvz = __fmul_rn(__fmul_rn(a.z, dt), 0.5f);
    v.x = __fadd_rn(v.x, dvx); v.y = __fadd_rn(v.y, dvy); v.z = __fadd_rn(v.z, dvz);
    p.x = fmaf(v.x, dt, p.x); p.y = fmaf(v.y, dt, p.y); p.z = fmaf(v.z, dt, p.z);
    v.x = __fadd_rn(v.x, dvx); v.y = __fadd_rn(v.y, dvy); v.z = __fadd_rn(v.z, dvz);
    node4[i] = p;
    velacc[2 * (size_t)i] = v;
}

// ---- host-boundary helpers: pack uploads, export logical buffers -------------------
__global__ void pack_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                            const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                            const float *__restrict__ mass, float4 *__restrict__ node4, float4 *__restrict__ velacc,
                            int *__restrict__ sorted, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    node4[i] = make_float4(x[i], y[i], z[i], mass[i]);
    velacc[2 * (size_t)i] = make_float4(vx[i], vy[i], vz[i], 0.0f);
    velacc[2 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    sorted[i] = 0;
}

// comp 0..3 of node4 for i < len
__global__ void export_node_kernel(const float4 *__restrict__ node4, int comp, float *__restrict__ out, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len) return;
    const float4 v = node4[i];
    out[i] = comp == 0 ? v.x : comp == 1 ? v.y : comp == 2 ? v.z : v.w;
}

// which = 0 (vel) or 1 (acc); zeros beyond the bodies (the reference never writes there)
__global__ void export_velacc_kernel(const float4 *__restrict__ velacc, int which, int comp, float *__restrict__ out,
                                     int n, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len) return;
    float r = 0.0f;
    if (i < n) {
        const float4 v = velacc[2 * (size_t)i + which];
        r = comp == 0 ? v.x : comp == 1 ? v.y : v.z;
    }
    out[i] = r;
}

// logical int arrays that exist only for cells: zeros for the first `skip` entries
__global__ void export_shifted_kernel(const int *__restrict__ src, long long skip, int *__restrict__ out, long long len) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= len) return;
    out[i] = i < skip ? 0 : src[i - skip];
}

// copyvertices.cl:14-17
__global__ void copy_vertices_kernel(const float4 *__restrict__ node4, const float4 *__restrict__ velacc,
                                     float4 *__restrict__ pos, float4 *__restrict__ vel, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pos) {
        const float4 p = node4[i];
        pos[i] = make_float4(p.x, p.y, p.z, 1.0f);
    }
    if (vel) {
        const float4 v = velacc[2 * (size_t)i];
        vel[i] = make_float4(v.x, v.y, v.z, 1.0f);
    }
}

// ---- seeded universe generators on the device (ch.fhnw.woipv.nbody.simulation.universe.*) -----------------------
// The reference's generators draw from the unseeded Math.random(); these draw from Philox4x32-10, one
// subsequence per body, so a universe is a pure function of (kind, seed, n) and needs no host memory.
// kind 0: RandomCubicUniverseGenerator.java:13-17   (U-0.5)*range per axis, v = 0, m = 1/n          (p0 = range)
// kind 1: PlummerUniverseGenerator.java:8-41        Plummer sphere, mass-fraction cut 0.999, m = 1/n
// kind 2: RotatingDiskGalaxyGenerator.java:17-43    disk of radius p0, velocity multiplier p1, body 0 = centre mass p2
__device__ __forceinline__ double philox_uniform(curandStatePhilox4_32_10_t *st) { return 1.0 - curand_uniform_double(st); }  // [0,1)

__global__ void __launch_bounds__(256) generate_kernel(float4 *__restrict__ node4, float4 *__restrict__ velacc,
                                                       int *__restrict__ sorted, int n, int kind, unsigned long long seed,
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
    node4[i] = make_float4(x, y, z, mass);
    velacc[2 * (size_t)i] = make_float4(vx, vy, vz, 0.0f);
    velacc[2 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    sorted[i] = 0;
}

// ---- diagnostics (GPUBH:305-365 printEnergy / printImpulse, on the device) ---------------------------
// out[0] = sum 1/2 m v^2, out[1..3] = sum m v, out[4] = sum m; double accumulation.
__global__ void __launch_bounds__(256) kinetic_kernel(const float4 *__restrict__ node4, const float4 *__restrict__ velacc,
                                                      double *__restrict__ out, int n) {
    double e = 0.0, px = 0.0, py = 0.0, pz = 0.0, ms = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float m = node4[i].w;
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
__global__ void __launch_bounds__(kPotTile) potential_kernel(const float4 *__restrict__ node4, double *__restrict__ out, int n,
                                                             float eps) {
    __shared__ float4 tile[kPotTile];
    const int i = blockIdx.x * kPotTile + threadIdx.x;
    const float4 pi = node4[min(i, n - 1)];
    double phi = 0.0;
    for (int j0 = 0; j0 < n; j0 += kPotTile) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < n ? node4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
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

}  // namespace bh
