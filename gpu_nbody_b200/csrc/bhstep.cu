// bhstep.cu -- host side of libbhstep.so: the C ABI declared in include/bhstep.h.
//
// Replaces the buffer / kernel / queue handling of
// src/ch/fhnw/woipv/nbody/simulation/gpu/GPUBarnesHutNBodySimulation.java (GPUBH)
// for the Barnes-Hut step.  No CPU fallback: without a CUDA device bh_create fails.
#include "bh_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "bhstep.h"

namespace {

constexpr int kProfSteps = 64;   // steps per bh_step call that get per-stage events
constexpr size_t kAccPad = 2048; // slack of the tree-order acceleration buffer: equal, 32-aligned slices for up to 64 ranks
constexpr int kNumTimers = BH_NUM_STAGES + 1;  // the six stages + the peer barrier

thread_local std::string g_createError;

struct Sim {
    int device = 0;
    int n = 0, m = 0, nc = 0;  // bodies, NUMBER_OF_NODES, cell slots (m - n + 1)
    float theta = 0.5f, thetaMacro = 0.25f, eps = 0.0025f, dt = 0.025f;
    int vote = 16;
    int numSMs = 0;
    cudaStream_t stream = nullptr, ownStream = nullptr;
    // device state; body arrays are double buffered (bodies are stored in tree order, see bh_kernels.cuh)
    float4 *body4[2] = {}, *velacc[2] = {}, *cell4 = nullptr, *octet = nullptr;
    char *accAlloc = nullptr;  // float4 acc[2][n + kAccPad] + peer flags + barrier sequence number (one IPC-exportable allocation)
    float4 *acc = nullptr;
    unsigned long long *flags = nullptr, *seq = nullptr;
    int2 *ometa = nullptr;
    int *child = nullptr, *start = nullptr, *count = nullptr, *perm = nullptr, *meta = nullptr, *parent = nullptr, *arrived = nullptr;
    int *spill = nullptr;  // the walk's stack spill area: kWalkSpillCap entries per group slot of every persistent CTA
    float *partials = nullptr;
    bh::Scalars *sc = nullptr;
    bh::Scalars *hostSc = nullptr;  // pinned mirror
    void *staging = nullptr;
    size_t stagingBytes = 0;
    // asynchronous vertex read-back (bh_copy_vertices_async): own staging, own stream, so that the device -> host copy of one
    // step's vertices overlaps the next upload (other PCIe direction) and the next step
    float4 *vtxStaging = nullptr;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evExported = nullptr, evCopied = nullptr;
    bool copyPending = false;
    float *deferredPos = nullptr, *deferredVel = nullptr;  // exported, but the device -> host copies are not enqueued yet
    bool copyDeferred = false;
    // asynchronous upload (bh_upload_async): the velocities travel on a second stream while the step's tree stages and walk run
    cudaStream_t upStream = nullptr;
    cudaEvent_t evVelReady = nullptr, evPosArrived = nullptr, evStateFree = nullptr, evPosPacked = nullptr;
    bool velPending = false;
    bool placed = false;     // the body buffers hold a complete state: every slot has a body with a distinct host number
    int *slotOf = nullptr;   // slot of the body with host number i (upload placement, vertex export by gather)
    bool slotOfValid = false;  // ... for the current state (any pass that moves bodies invalidates it)
    bool stagingBusy = false;       // an asynchronous upload's pack kernels may still be reading the staging buffer
    bool stagingSharedUse = false;  // the simulation's stream has used the staging buffer (bh_read ...) since the last upload
    int cur = 0;           // buffers holding the current body state
    int treePhase = 0;     // buffers the tree (child[]) was built from
    unsigned stagesRun = 0;  // bit per stage that has run since the upload: a stage needs the outputs of the ones before it
    bool havePerm = false;   // a sort has run since the upload
    bool permValid = false;  // perm[] refers to slots of the current buffers (false: bodies already lie in tree order)
    // launch geometry
    int bboxGrid = 1, buildGrid = 1, summGrid = 1, sortGrid = 1, walkGrid = 1;
    // options
    bool profiling = false, counting = false, forceDeep = false;
    int insertionOrder = 1;
    float4 *vertexPos = nullptr, *vertexVel = nullptr;  // fused copyVertices destinations (device pointers)
    // profiling
    cudaEvent_t ev[kProfSteps][kNumTimers + 1] = {};
    bool evCreated = false;
    int evSteps = 0;
    double stageMs[kNumTimers] = {};
    int64_t stageLaunches[BH_NUM_STAGES] = {};
    int64_t stepsTimed = 0;
    // multi-GPU: the rank's slice of the tree order; peers' acceleration buffers and flags mapped through CUDA IPC
    int sliceFirst = 0, sliceCount = 0;
    int nranks = 1, rank = 0;
    bool p2p = false;
    char *peerAlloc[bh::kMaxPeers] = {};
    // CUDA graphs of one step (one per parity of the body buffers)
    bool useGraph = true;
    cudaGraphExec_t graphExec[2] = {};
    cudaStream_t graphStream = nullptr;
    std::string lastError;
};

int fail(Sim *s, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) s->lastError = buf; else g_createError = buf;
    return code;
}

#define BH_CUDA(s, call)                                                                               \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail((s), BH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline Sim *S(bh_sim *p) { return reinterpret_cast<Sim *>(p); }
inline size_t accStride(const Sim *s) { return (size_t)s->n + kAccPad; }
inline size_t accBytes(const Sim *s) { return sizeof(float4) * 2 * accStride(s); }

void dropGraphs(Sim *s) {
    for (auto &g : s->graphExec)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}

int ensureStaging(Sim *s, size_t bytes) {
    if (s->stagingBusy) {  // an asynchronous upload's pack kernels may still be reading it
        BH_CUDA(s, cudaStreamWaitEvent(s->stream, s->evPosPacked, 0));
        s->stagingBusy = false;
    }
    s->stagingSharedUse = true;
    if (bytes <= s->stagingBytes) return BH_OK;
    if (s->staging) cudaFree(s->staging);
    s->staging = nullptr;
    s->stagingBytes = 0;
    BH_CUDA(s, cudaMalloc(&s->staging, bytes));
    s->stagingBytes = bytes;
    return BH_OK;
}

// An asynchronous upload may still be delivering the velocities: everything that touches them waits for it (stream order)
int settleVel(Sim *s) {
    if (s->velPending) {
        BH_CUDA(s, cudaStreamWaitEvent(s->stream, s->evVelReady, 0));
        s->velPending = false;
    }
    return BH_OK;
}

// slotOf[] for the current state (the caller has settled the velocities: the host numbers live beside them)
int ensureSlotOf(Sim *s) {
    if (!s->slotOf) BH_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->slotOf), sizeof(int) * (size_t)s->n));
    if (!s->slotOfValid) {
        bh::slot_of_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->velacc[s->cur], s->slotOf, s->n);
        BH_CUDA(s, cudaGetLastError());
        s->slotOfValid = true;
    }
    return BH_OK;
}

// copyvertices.cl on the current state, host numbering, to device memory
int exportVertices(Sim *s, float4 *pos, float4 *vel) {
    if (!s->placed) {  // nothing has been uploaded or generated yet: slot order (all host numbers are zero)
        bh::copy_vertices_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->body4[s->cur], s->velacc[s->cur], pos, vel, s->n);
    } else {
        int rc = ensureSlotOf(s);
        if (rc) return rc;
        bh::copy_vertices_gather_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->body4[s->cur], s->velacc[s->cur], s->slotOf, pos, vel, s->n);
    }
    BH_CUDA(s, cudaGetLastError());
    return BH_OK;
}

// bh_copy_vertices_async exports at once but enqueues its device -> host copies lazily: if an upload follows, they start
// behind that upload's position copies (which gate the next step) instead of competing with them for the link
int flushCopy(Sim *s, cudaEvent_t after) {
    if (!s->copyDeferred) return BH_OK;
    const size_t bytes = sizeof(float4) * (size_t)s->n;
    BH_CUDA(s, cudaStreamWaitEvent(s->copyStream, s->evExported, 0));
    if (after) BH_CUDA(s, cudaStreamWaitEvent(s->copyStream, after, 0));
    if (s->deferredPos) BH_CUDA(s, cudaMemcpyAsync(s->deferredPos, s->vtxStaging, bytes, cudaMemcpyDeviceToHost, s->copyStream));
    if (s->deferredVel) BH_CUDA(s, cudaMemcpyAsync(s->deferredVel, s->vtxStaging + s->n, bytes, cudaMemcpyDeviceToHost, s->copyStream));
    BH_CUDA(s, cudaEventRecord(s->evCopied, s->copyStream));
    s->copyDeferred = false;
    s->copyPending = true;
    return BH_OK;
}

int resetState(Sim *s, bool everything = false) {
    // GPUBH:155-179: everything zero except step = -1, maxDepth = 1.  The tree buffers (1.1 GB at 10^7 bodies) are
    // zeroed as a whole once, at creation; later resets clear the cells that builds have allocated since (the device
    // knows which: Scalars::lowWater), a third of the buffers in a typical run.
    if (everything) {
        BH_CUDA(s, cudaMemsetAsync(s->child, 0, sizeof(int) * 8 * (size_t)s->nc, s->stream));
        BH_CUDA(s, cudaMemsetAsync(s->start, 0, sizeof(int) * (size_t)s->nc, s->stream));
        BH_CUDA(s, cudaMemsetAsync(s->count, 0, sizeof(int) * (size_t)s->nc, s->stream));
        BH_CUDA(s, cudaMemsetAsync(s->cell4, 0, sizeof(float4) * (size_t)s->nc, s->stream));
    } else {
        bh::clear_tree_kernel<<<s->numSMs * 8, 256, 0, s->stream>>>(s->child, s->start, s->count, s->cell4, s->sc, s->n, s->m);
        BH_CUDA(s, cudaGetLastError());
    }
    bh::init_scalars_kernel<<<1, 1, 0, s->stream>>>(s->sc);
    BH_CUDA(s, cudaGetLastError());
    s->cur = 0;
    s->treePhase = 0;
    s->stagesRun = 0;
    s->havePerm = false;
    s->permValid = false;
    return BH_OK;
}

bh::PeerBuffers accDestinations(const Sim *s, bool peers) {
    bh::PeerBuffers dst;
    dst.count = 1;
    dst.buf[0] = s->acc;
    dst.phaseStride = s->p2p ? (unsigned)accStride(s) : 0u;
    if (peers)
        for (int r = 0; r < s->nranks; ++r)
            if (r != s->rank) dst.buf[dst.count++] = reinterpret_cast<float4 *>(s->peerAlloc[r]);
    return dst;
}

// force walk for tree-order slots [first, first+cnt) into the acceleration buffer(s)
void launchWalk(Sim *s, int first, int cnt, bool peers, bool potential = false) {
    if (cnt <= 0) return;
    const bh::PeerBuffers dst = accDestinations(s, peers);
    const int *perm = s->permValid ? s->perm : nullptr;
    const float4 *body = s->body4[s->cur];
    if (s->vote == 16 && !s->forceDeep) {
        // persistent CTAs; their warps draw chunks of eight vote groups from a ticket in the scalars
        const int groups = (cnt + 15) / 16;
        const int grid = std::max(1, std::min(s->walkGrid, (groups + 31) / 32));
        const size_t smem = sizeof(bh::WalkShared);
        bh::reset_ticket_kernel<<<1, 1, 0, s->stream>>>(s->sc);
        if (potential)
            bh::walk_kernel<false, true><<<grid, bh::kWalkThreads, smem, s->stream>>>(body, s->octet, s->ometa, perm, dst, s->sc, s->spill, s->n, first, cnt, s->eps);
        else if (s->counting)
            bh::walk_kernel<true, false><<<grid, bh::kWalkThreads, smem, s->stream>>>(body, s->octet, s->ometa, perm, dst, s->sc, s->spill, s->n, first, cnt, s->eps);
        else
            bh::walk_kernel<false, false><<<grid, bh::kWalkThreads, smem, s->stream>>>(body, s->octet, s->ometa, perm, dst, s->sc, s->spill, s->n, first, cnt, s->eps);
        return;
    }
    // 32-wide votes (not reference-exact), or the shared-stack walk on request
    const int chunks = (cnt + bh::kForce2Bodies - 1) / bh::kForce2Bodies;
#define BH_DEEP(V, C) bh::deep_walk_kernel<V, C><<<chunks, bh::kForce2Threads, 0, s->stream>>>(body, s->octet, s->ometa, s->meta, perm, dst, s->sc, s->n, s->m, first, cnt, s->thetaMacro, s->eps)
    if (s->vote == 16) { if (s->counting) BH_DEEP(16, true); else BH_DEEP(16, false); }
    else { if (s->counting) BH_DEEP(32, true); else BH_DEEP(32, false); }
#undef BH_DEEP
}

int launchBarrier(Sim *s) {
    bh::PeerFlags pf;
    pf.count = s->nranks;
    pf.rank = s->rank;
    for (int r = 0; r < s->nranks; ++r)
        pf.flags[r] = reinterpret_cast<unsigned long long *>((r == s->rank ? s->accAlloc : s->peerAlloc[r]) + accBytes(s));
    bh::barrier_kernel<<<1, 32, 0, s->stream>>>(pf, s->seq, s->sc, 8000000000ll);  // ~4 s at 1.9 GHz
    BH_CUDA(s, cudaGetLastError());
    return BH_OK;
}

// integrate (+ velocity correction when apply) (+ move every body to its tree-order slot when a fresh sort exists)
int launchFinish(Sim *s, bool apply) {
    const bool permute = s->insertionOrder == 1 && s->permValid;
    const int in = s->cur, out = permute ? in ^ 1 : in;
    const int *perm = s->permValid ? s->perm : nullptr;
    const unsigned stride = s->p2p ? (unsigned)accStride(s) : 0u;
    const int grid = (s->n + 255) / 256;
#define BH_FINISH(A, P)                                                                                                       \
    bh::finish_kernel<A, P><<<grid, 256, 0, s->stream>>>(s->body4[in], s->velacc[in], s->body4[out], s->velacc[out], s->acc, stride, \
                                                         perm, s->sc, s->n, s->dt, s->vertexPos, s->vertexVel)
    if (apply) { if (permute) BH_FINISH(true, true); else BH_FINISH(true, false); }
    else { if (permute) BH_FINISH(false, true); else BH_FINISH(false, false); }
#undef BH_FINISH
    BH_CUDA(s, cudaGetLastError());
    if (permute) {
        s->cur = out;
        s->slotOfValid = false;
        s->permValid = false;  // the bodies now lie in tree order
        s->stagesRun &= ~(1u << BH_STAGE_BUILD);  // ... and child[] names them by their old slots: no summarise / sort without a rebuild
    }
    return BH_OK;
}

int launchSort(Sim *s) {
    // cooperative launch: every CTA is resident (the kernel's waits are on cells processed by resident threads)
    const int *child = s->child, *count = s->count;
    int *start = s->start, *perm = s->perm;
    bh::Scalars *sc = s->sc;
    int n = s->n, m = s->m;
    void *args[] = {&child, &count, &start, &perm, &sc, &n, &m};
    BH_CUDA(s, cudaLaunchCooperativeKernel(reinterpret_cast<void *>(bh::sort_kernel), dim3(s->sortGrid), dim3(bh::kSortThreads), args, 0, s->stream));
    s->havePerm = true;
    s->permValid = true;
    return BH_OK;
}

// `fused`: inside bh_step the force stage leaves the accelerations in the tree-order buffer and the integrate stage
// applies them (one pass over the bodies); as single stages each completes its own reference semantics.
int launchStage(Sim *s, int stage, bool fused) {
    const int n = s->n, m = s->m;
    // The reference would run any kernel on whatever its buffers hold; here the tree buffers are uninitialised device
    // memory until the stages before have run since the upload, and the tree names bodies by slots that a reordering
    // finish pass changes, so calls that would read such state are refused.  (A force walk over a stale but complete
    // tree is allowed, as in the reference: it reads the cells' walk records only.)
    constexpr unsigned B = 1u << BH_STAGE_BBOX, T = 1u << BH_STAGE_BUILD, U = 1u << BH_STAGE_SUMMARIZE, O = 1u << BH_STAGE_SORT;
    static const unsigned needs[BH_NUM_STAGES] = {0u, B, T, T | U, U | O, 0u};
    static const char *const names[BH_NUM_STAGES] = {"bounding_box", "build_tree", "summarize", "sort", "calculate_force", "integrate"};
    if (stage >= 0 && stage < BH_NUM_STAGES && (s->stagesRun & needs[stage]) != needs[stage])
        return fail(s, BH_ERR_ARG, "%s called before the stages it depends on have run (since the upload / the last reordering)", names[stage]);
    if (stage == BH_STAGE_BBOX) s->stagesRun &= ~(T | U | O);  // the root is reset: the tree is being rebuilt
    if (stage == BH_STAGE_BUILD) s->stagesRun &= ~(U | O);
    if (stage >= 0 && stage < BH_NUM_STAGES) s->stagesRun |= 1u << stage;
    switch (stage) {
    case BH_STAGE_BBOX:
        bh::bbox_kernel<<<s->bboxGrid, bh::kBboxThreads, 0, s->stream>>>(s->body4[s->cur], s->cell4, s->child, s->start, s->count,
                                                                         s->arrived, s->partials, s->sc, n, m);
        break;
    case BH_STAGE_BUILD:
        bh::build_kernel<<<s->buildGrid, bh::kBuildThreads, 0, s->stream>>>(s->body4[s->cur], s->cell4, s->child, s->start, s->count,
                                                                            s->parent, s->arrived, s->sc, n, m);
        s->treePhase = s->cur;
        break;
    case BH_STAGE_SUMMARIZE:
        bh::summarize_kernel<<<s->summGrid, bh::kSummThreads, 0, s->stream>>>(s->body4[s->cur], s->cell4, s->child, s->octet, s->ometa,
                                                                              s->meta, s->count, s->parent, s->arrived, s->sc, n, m,
                                                                              s->thetaMacro, s->eps);
        break;
    case BH_STAGE_SORT: {
        int rc = launchSort(s);
        if (rc) return rc;
        break;
    }
    case BH_STAGE_FORCE: {
        if (s->counting) bh::reset_counters_kernel<<<1, 1, 0, s->stream>>>(s->sc);
        const bool sliced = fused && s->p2p;
        launchWalk(s, sliced ? s->sliceFirst : 0, sliced ? s->sliceCount : n, sliced);
        if (!fused) {
            int rc = settleVel(s);
            if (rc) return rc;
            const unsigned stride = s->p2p ? (unsigned)accStride(s) : 0u;
            bh::apply_acc_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->acc, stride, s->permValid ? s->perm : nullptr,
                                                                       s->velacc[s->cur], s->sc, n, s->dt);
        }
        break;
    }
    case BH_STAGE_INTEGRATE: {
        int rc = settleVel(s);
        if (rc) return rc;
        rc = launchFinish(s, fused);
        if (rc) return rc;
        break;
    }
    default:
        return fail(s, BH_ERR_ARG, "unknown stage %d", stage);
    }
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[stage] += 1;
    if (stage == BH_STAGE_FORCE && !fused) s->stageLaunches[stage]++;
    return BH_OK;
}

// sync, pull the scalars, report the device error buffer
int finish(Sim *s) {
    BH_CUDA(s, cudaMemcpyAsync(s->hostSc, s->sc, sizeof(bh::Scalars), cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    if (s->profiling && s->evSteps > 0) {
        for (int i = 0; i < s->evSteps; ++i)
            for (int st = 0; st < kNumTimers; ++st) {
                float ms = 0.0f;
                if (cudaEventElapsedTime(&ms, s->ev[i][st], s->ev[i][st + 1]) == cudaSuccess) s->stageMs[st] += ms;
            }
        s->stepsTimed += s->evSteps;
        s->evSteps = 0;
    }
    if (s->hostSc->error != 0) {
        const int e = s->hostSc->error;
        fail(s, e, "device error buffer = %d (%s)", e,
             e == 1 ? "cell pool exhausted or tree deeper than 64 levels"
                    : e == 2 ? "device-side wait exceeded its spin budget" : "peer barrier timed out");
        return e;
    }
    return BH_OK;
}

// The launches of one step in order; timers[] (optional) = kNumTimers + 1 events recorded between them in the order
// bbox, build, summarise, sort, force, barrier, finish.
int launchStep(Sim *s, cudaEvent_t *timers) {
    static const int order[kNumTimers] = {BH_STAGE_BBOX, BH_STAGE_BUILD, BH_STAGE_SUMMARIZE, BH_STAGE_SORT, BH_STAGE_FORCE, -1,
                                          BH_STAGE_INTEGRATE};
    // timer slots: stage index for the six stages, kTimerBarrier for the barrier; events are recorded in launch order
    for (int k = 0; k < kNumTimers; ++k) {
        if (timers) BH_CUDA(s, cudaEventRecord(timers[k], s->stream));
        if (order[k] < 0) {
            if (s->p2p) {
                int rc = launchBarrier(s);
                if (rc) return rc;
                s->stageLaunches[BH_STAGE_FORCE]++;
            }
            continue;
        }
        int rc = launchStage(s, order[k], true);
        if (rc) return rc;
    }
    if (timers) BH_CUDA(s, cudaEventRecord(timers[kNumTimers], s->stream));
    return BH_OK;
}

// One step = a fixed sequence of launches with fixed arguments per parity of the body buffers: captured once into a
// CUDA graph per parity and replayed (one launch call per step; matters for small universes such as the reference's
// 32 768-body default, where a step is a few hundred microseconds, and for the multi-GPU step, whose cross-rank
// barrier is a kernel of its own).  Not used while per-stage events or counters are on.
int graphStep(Sim *s) {
    if (s->graphStream != s->stream) { dropGraphs(s); s->graphStream = s->stream; }
    const int parity = s->cur;
    const bool willPermute = s->insertionOrder == 1;
    if (!s->graphExec[parity]) {
        cudaGraph_t graph = nullptr;
        const int cur0 = s->cur, tree0 = s->treePhase;
        const bool have0 = s->havePerm, valid0 = s->permValid;
        const unsigned stages0 = s->stagesRun;
        int64_t launches0[BH_NUM_STAGES];
        memcpy(launches0, s->stageLaunches, sizeof launches0);
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();    // e.g. a stream that cannot be captured: plain launches from now on
            s->useGraph = false;
            return BH_ERR_ARG;     // tells stepAsync to launch this step the ordinary way
        }
        const int rc = launchStep(s, nullptr);
        const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        // the capture did not run anything: restore the host-side state
        s->cur = cur0; s->treePhase = tree0; s->havePerm = have0; s->permValid = valid0; s->stagesRun = stages0;
        memcpy(s->stageLaunches, launches0, sizeof launches0);
        if (rc != BH_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            s->useGraph = false;   // e.g. a driver that cannot capture a cooperative launch: plain launches from now on
            return BH_ERR_ARG;
        }
        const cudaError_t ei = cudaGraphInstantiate(&s->graphExec[parity], graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            s->graphExec[parity] = nullptr;
            cudaGetLastError();
            s->useGraph = false;
            return BH_ERR_ARG;
        }
    }
    BH_CUDA(s, cudaGraphLaunch(s->graphExec[parity], s->stream));
    // host-side mirror of what the replayed launches did
    s->stagesRun = ((1u << BH_NUM_STAGES) - 1u) & ~(willPermute ? (1u << BH_STAGE_BUILD) : 0u);
    s->treePhase = parity;
    s->havePerm = true;
    if (willPermute) { s->cur = parity ^ 1; s->permValid = false; s->slotOfValid = false; } else { s->permValid = true; }
    for (int st = 0; st < BH_NUM_STAGES; ++st) s->stageLaunches[st] += 1;
    if (s->p2p) s->stageLaunches[BH_STAGE_FORCE]++;  // the peer barrier
    return BH_OK;
}

int stepAsync(Sim *s, int nsteps) {
    int rcf = flushCopy(s, nullptr);
    if (rcf) return rcf;
    for (int i = 0; i < nsteps; ++i) {
        // (the legacy default stream cannot be captured)
        if (s->useGraph && !s->profiling && !s->counting && !s->velPending && s->stream != nullptr) {
            int rc = graphStep(s);
            if (rc == BH_OK) continue;
            if (s->useGraph) return rc;  // a real failure; otherwise capture was refused and the graph is now off
        }
        const bool prof = s->profiling && s->evSteps < kProfSteps;
        if (prof && !s->evCreated) {
            for (auto &row : s->ev)
                for (auto &e : row) BH_CUDA(s, cudaEventCreate(&e));
            s->evCreated = true;
        }
        int rc = launchStep(s, prof ? s->ev[s->evSteps] : nullptr);
        if (rc) return rc;
        if (prof) s->evSteps++;
    }
    return BH_OK;
}

int singleStage(Sim *s, int stage) {
    if (s->profiling) {
        // single stages are timed through one-off events folded into the same accumulators
        cudaEvent_t a, b;
        BH_CUDA(s, cudaEventCreate(&a));
        BH_CUDA(s, cudaEventCreate(&b));
        BH_CUDA(s, cudaEventRecord(a, s->stream));
        int rc = launchStage(s, stage, false);
        if (rc) return rc;
        BH_CUDA(s, cudaEventRecord(b, s->stream));
        rc = finish(s);
        float ms = 0.0f;
        // timer slots follow the launch order of a step: 0-4 = bbox..force, 5 = peer barrier, 6 = finish
        if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) s->stageMs[stage == BH_STAGE_INTEGRATE ? kNumTimers - 1 : stage] += ms;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        return rc;
    }
    int rc = launchStage(s, stage, false);
    if (rc) return rc;
    return finish(s);
}

bool validSlice(const Sim *s, int32_t first, int32_t count) {
    return first >= 0 && count >= 0 && (int64_t)first + count <= s->n && (first % s->vote) == 0;
}

}  // namespace

extern "C" {

int32_t bh_abi_version(void) { return 2; }

int32_t bh_number_of_nodes(int32_t nbodies) {
    // GPUBH:219-227 with maxComputeUnits = 16 (GPUBH:126) and WARPSIZE = 16 (GPUBH:47)
    int64_t nodes = (int64_t)nbodies * 2;
    if (nodes < 1024 * 16) nodes = 1024 * 16;
    while ((nodes & 15) != 0) ++nodes;
    return nodes > INT32_MAX - 1 ? -1 : (int32_t)nodes;
}

const char *bh_last_error(bh_sim *sim) { return sim ? S(sim)->lastError.c_str() : g_createError.c_str(); }

int bh_create(bh_sim **out, int32_t nbodies, float theta, float eps2, float dt, int32_t vote_width, int32_t device) {
    if (!out) return fail(nullptr, BH_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (nbodies < 1) return fail(nullptr, BH_ERR_ARG, "nbodies must be >= 1");
    if (vote_width != 16 && vote_width != 32) return fail(nullptr, BH_ERR_ARG, "vote_width must be 16 or 32");
    const int32_t m = bh_number_of_nodes(nbodies);
    // child rows are addressed as 8*(cell-N) in size_t; node indices must fit int32, walk entries 27 bits
    if (m < 0 || (int64_t)m - nbodies + 1 > (int64_t)bh::kEntryMask) return fail(nullptr, BH_ERR_ARG, "nbodies too large (at most 2^27 - 2 cells)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, BH_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, BH_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    Sim *s = new (std::nothrow) Sim();
    if (!s) return fail(nullptr, BH_ERR_ALLOC, "out of host memory");
    s->device = device;
    s->n = nbodies;
    s->m = m;
    s->nc = m - nbodies + 1;
    s->theta = theta;
    s->thetaMacro = theta * theta;
    s->eps = eps2;
    s->dt = dt;
    s->vote = vote_width;
    s->sliceFirst = 0;
    s->sliceCount = nbodies;
    auto bail = [&](int code, const char *what, cudaError_t e) {
        fail(nullptr, code, "%s: %s", what, cudaGetErrorString(e));
        bh_destroy(reinterpret_cast<bh_sim *>(s));
        return code;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(BH_ERR_CUDA, "cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(BH_ERR_CUDA, "cudaGetDeviceProperties", e);
    s->numSMs = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) return bail(BH_ERR_CUDA, "device cannot launch cooperative kernels", cudaErrorNotSupported);
    if ((e = cudaStreamCreateWithFlags(&s->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return bail(BH_ERR_CUDA, "cudaStreamCreate", e);
    s->stream = s->ownStream;
    const size_t n = nbodies, nc = s->nc;
#define BH_ALLOC(ptr, bytes) \
    if ((e = cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes))) != cudaSuccess) return bail(BH_ERR_ALLOC, "cudaMalloc " #ptr, e)
    for (int b = 0; b < 2; ++b) {
        BH_ALLOC(s->body4[b], sizeof(float4) * n);
        BH_ALLOC(s->velacc[b], sizeof(float4) * 2 * n);
    }
    BH_ALLOC(s->cell4, sizeof(float4) * nc);
    BH_ALLOC(s->octet, sizeof(float4) * 8 * nc);
    BH_ALLOC(s->ometa, sizeof(int2) * 8 * nc);
    BH_ALLOC(s->accAlloc, accBytes(s) + 512);  // two phases, peer flags, barrier sequence number
    BH_ALLOC(s->child, sizeof(int) * 8 * nc);
    BH_ALLOC(s->start, sizeof(int) * nc);
    BH_ALLOC(s->count, sizeof(int) * nc);
    BH_ALLOC(s->meta, sizeof(int) * nc);
    BH_ALLOC(s->parent, sizeof(int) * nc);
    BH_ALLOC(s->arrived, sizeof(int) * nc);
    BH_ALLOC(s->perm, sizeof(int) * n);
    BH_ALLOC(s->sc, sizeof(bh::Scalars));
#undef BH_ALLOC
    s->acc = reinterpret_cast<float4 *>(s->accAlloc);
    s->flags = reinterpret_cast<unsigned long long *>(s->accAlloc + accBytes(s));
    s->seq = s->flags + bh::kMaxPeers;
    if ((e = cudaMallocHost(reinterpret_cast<void **>(&s->hostSc), sizeof(bh::Scalars))) != cudaSuccess) return bail(BH_ERR_ALLOC, "cudaMallocHost", e);
    // launch geometry: streaming kernels a few CTAs per SM; sort (which waits on other threads) exactly as many
    // CTAs as are resident at once (and it is launched cooperatively).
    int perSM = 1;
    s->bboxGrid = (int)std::min<size_t>((n + bh::kBboxThreads - 1) / bh::kBboxThreads, (size_t)s->numSMs * 4);
    if ((e = cudaMalloc(reinterpret_cast<void **>(&s->partials), sizeof(float) * 6 * s->bboxGrid)) != cudaSuccess) return bail(BH_ERR_ALLOC, "cudaMalloc partials", e);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, bh::build_kernel, bh::kBuildThreads, 0);
    // twice the resident CTAs: shorter runs per lane and a second wave that evens out the lanes' very unequal run times
    // (measured at 10^7 bodies: 1x 1.34 ms, 2x 1.14 ms, 4x 1.16 ms, 8x 1.40 ms; summarise gains 7 % from the finer cell order)
    s->buildGrid = (int)std::min<size_t>((n + bh::kBuildThreads - 1) / bh::kBuildThreads, (size_t)s->numSMs * std::max(perSM, 1) * 2);
    s->summGrid = (int)std::min<size_t>((nc + bh::kSummThreads - 1) / bh::kSummThreads, (size_t)s->numSMs * 64);  // never waits: any grid
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, bh::sort_kernel, bh::kSortThreads, 0);
    s->sortGrid = s->numSMs * std::max(perSM, 1);
    // tiny problems: do not launch more waiting threads than there can be cells
    const int cellBlocks = (int)((nc + bh::kSummThreads - 1) / bh::kSummThreads);
    s->sortGrid = std::max(1, std::min(s->sortGrid, cellBlocks));
    cudaFuncSetAttribute(bh::walk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(bh::WalkShared));
    cudaFuncSetAttribute(bh::walk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(bh::WalkShared));
    cudaFuncSetAttribute(bh::walk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(bh::WalkShared));
    perSM = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, bh::walk_kernel<false, false>, bh::kWalkThreads, sizeof(bh::WalkShared));
    s->walkGrid = s->numSMs * std::max(1, std::min(perSM, bh::kWalkCtasPerSM));
    if ((e = cudaMalloc(reinterpret_cast<void **>(&s->spill), sizeof(int) * (size_t)s->walkGrid * bh::kWalkWarps * bh::kWalkGroups * bh::kWalkSpillCap)) != cudaSuccess)
        return bail(BH_ERR_ALLOC, "cudaMalloc spill", e);
    for (int b = 0; b < 2; ++b) {
        if ((e = cudaMemsetAsync(s->body4[b], 0, sizeof(float4) * n, s->stream)) != cudaSuccess) return bail(BH_ERR_CUDA, "cudaMemset", e);
        cudaMemsetAsync(s->velacc[b], 0, sizeof(float4) * 2 * n, s->stream);
    }
    cudaMemsetAsync(s->perm, 0, sizeof(int) * n, s->stream);
    cudaMemsetAsync(s->accAlloc, 0, accBytes(s) + 512, s->stream);
    if (resetState(s, true) != BH_OK || cudaStreamSynchronize(s->stream) != cudaSuccess) {
        g_createError = s->lastError.empty() ? "initial reset failed" : s->lastError;
        bh_destroy(reinterpret_cast<bh_sim *>(s));
        return BH_ERR_CUDA;
    }
    *out = reinterpret_cast<bh_sim *>(s);
    return BH_OK;
}

static void closePeers(Sim *s) {
    for (int r = 0; r < bh::kMaxPeers; ++r) {
        if (s->peerAlloc[r] && s->peerAlloc[r] != s->accAlloc) cudaIpcCloseMemHandle(s->peerAlloc[r]);
        s->peerAlloc[r] = nullptr;
    }
    s->p2p = false;
    s->nranks = 1;
    s->rank = 0;
}

void bh_destroy(bh_sim *sim) {
    if (!sim) return;
    Sim *s = S(sim);
    cudaSetDevice(s->device);
    flushCopy(s, nullptr);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copyStream) cudaStreamSynchronize(s->copyStream);
    if (s->upStream) cudaStreamSynchronize(s->upStream);
    dropGraphs(s);
    closePeers(s);
    for (int b = 0; b < 2; ++b) { cudaFree(s->body4[b]); cudaFree(s->velacc[b]); }
    cudaFree(s->cell4); cudaFree(s->octet); cudaFree(s->ometa); cudaFree(s->accAlloc);
    cudaFree(s->child); cudaFree(s->start); cudaFree(s->count); cudaFree(s->perm); cudaFree(s->meta); cudaFree(s->parent); cudaFree(s->arrived);
    cudaFree(s->partials); cudaFree(s->sc); cudaFree(s->staging); cudaFree(s->spill); cudaFree(s->slotOf);
    if (s->hostSc) cudaFreeHost(s->hostSc);
    if (s->evCreated)
        for (auto &row : s->ev)
            for (auto &e : row) cudaEventDestroy(e);
    cudaFree(s->vtxStaging);
    if (s->evExported) cudaEventDestroy(s->evExported);
    if (s->evCopied) cudaEventDestroy(s->evCopied);
    if (s->copyStream) cudaStreamDestroy(s->copyStream);
    if (s->evVelReady) cudaEventDestroy(s->evVelReady);
    if (s->evPosArrived) cudaEventDestroy(s->evPosArrived);
    if (s->evStateFree) cudaEventDestroy(s->evStateFree);
    if (s->evPosPacked) cudaEventDestroy(s->evPosPacked);
    if (s->upStream) cudaStreamDestroy(s->upStream);
    if (s->ownStream) cudaStreamDestroy(s->ownStream);
    delete s;
}

#define BH_ENTER(sim)                                          \
    if (!(sim)) return BH_ERR_ARG;                             \
    Sim *s = S(sim);                                           \
    BH_CUDA(s, cudaSetDevice(s->device))

#define BH_ENTER_SETTLED(sim)                                  \
    BH_ENTER(sim);                                             \
    do { int rcv_ = settleVel(s); if (rcv_) return rcv_; } while (0)

int bh_set_theta_macro(bh_sim *sim, float theta_macro) {
    BH_ENTER(sim);
    dropGraphs(s);  // kernel arguments change
    s->thetaMacro = theta_macro;
    return BH_OK;
}

int bh_set_stream(bh_sim *sim, void *cuda_stream) {
    BH_ENTER_SETTLED(sim);
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    s->stream = reinterpret_cast<cudaStream_t>(cuda_stream);  // NULL = CUDA's default stream, as everywhere in CUDA
    return BH_OK;
}

int bh_use_private_stream(bh_sim *sim) {
    BH_ENTER_SETTLED(sim);
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    s->stream = s->ownStream;
    return BH_OK;
}

int bh_set_profiling(bh_sim *sim, int32_t on) {
    BH_ENTER(sim);
    s->profiling = on != 0;
    return BH_OK;
}

int bh_set_counting(bh_sim *sim, int32_t on) {
    BH_ENTER(sim);
    s->counting = on != 0;
    return BH_OK;
}

int bh_set_graph(bh_sim *sim, int32_t on) {
    BH_ENTER(sim);
    s->useGraph = on != 0;
    return BH_OK;
}

int bh_set_insertion_order(bh_sim *sim, int32_t mode) {
    BH_ENTER(sim);
    if (mode != 0 && mode != 1) return fail(s, BH_ERR_ARG, "insertion order must be 0 or 1");
    if (mode != s->insertionOrder) dropGraphs(s);
    s->insertionOrder = mode;
    return BH_OK;
}

int bh_set_force_deep_walk(bh_sim *sim, int32_t on) {
    BH_ENTER(sim);
    if ((on != 0) != s->forceDeep) dropGraphs(s);
    s->forceDeep = on != 0;
    return BH_OK;
}

int bh_set_vertex_buffers(bh_sim *sim, void *pos4_device, void *vel4_device) {
    BH_ENTER(sim);
    dropGraphs(s);
    s->vertexPos = static_cast<float4 *>(pos4_device);
    s->vertexVel = static_cast<float4 *>(vel4_device);
    return BH_OK;
}

// async: every host -> device copy runs on a second stream, so that a caller that does not wait between steps
// (bh_step_async) gets the NEXT step's inputs across the link while the current step computes; the simulation's own
// stream only waits for the positions (before pack_pos) and, in the finish pass, for the velocities.  Order on the
// upload stream: [staging free] copy positions+masses -> evPosArrived -> copy velocities -> [state free] pack_vel
// -> evVelReady.  The call returns without waiting for anything.
static int uploadImpl(Sim *s, const float *const src[7], cudaMemcpyKind kind, bool async) {
    for (int i = 0; i < 7; ++i)
        if (!src[i]) return fail(s, BH_ERR_ARG, "NULL input array %d", i);
    int rc = BH_OK;
    const size_t n = s->n;
    const int grid = (s->n + 255) / 256;
    const float *dev[7];
    static const int order[7] = {0, 1, 2, 6, 3, 4, 5};  // positions and masses first
    if (kind != cudaMemcpyHostToDevice) async = false;
    if (!async) {
        rc = settleVel(s);  // a previous asynchronous upload still owns the staging buffer
        if (rc) return rc;
    }
    if (kind == cudaMemcpyHostToDevice) {
        if (async && !s->upStream) {
            BH_CUDA(s, cudaStreamCreateWithFlags(&s->upStream, cudaStreamNonBlocking));
            BH_CUDA(s, cudaEventCreateWithFlags(&s->evVelReady, cudaEventDisableTiming));
            BH_CUDA(s, cudaEventCreateWithFlags(&s->evStateFree, cudaEventDisableTiming));
            BH_CUDA(s, cudaEventCreateWithFlags(&s->evPosPacked, cudaEventDisableTiming));
        }
        if (!s->evPosArrived) BH_CUDA(s, cudaEventCreateWithFlags(&s->evPosArrived, cudaEventDisableTiming));
        if (s->stagingBytes < sizeof(float) * 7 * n) {  // (re)allocating the staging buffer: nothing may be using it
            BH_CUDA(s, cudaStreamSynchronize(s->stream));
            if (s->upStream) BH_CUDA(s, cudaStreamSynchronize(s->upStream));
            rc = ensureStaging(s, sizeof(float) * 7 * n);
            if (rc) return rc;
            s->stagingBusy = false;
        }
        float *stg = static_cast<float *>(s->staging);
        cudaStream_t up = async ? s->upStream : s->stream;
        if (async) {
            // the staging buffer is free once the previous upload's pack kernels have read it; other users of the
            // staging buffer (bh_read ...) run on the simulation's stream: everything enqueued there so far comes first
            // only as far as the STATE is concerned (evStateFree, awaited before pack_vel), not for the copies
            if (s->stagingBusy) BH_CUDA(s, cudaStreamWaitEvent(up, s->evPosPacked, 0));
            if (s->stagingSharedUse) {  // a bh_read / dump used the staging buffer since: order the copies behind it
                BH_CUDA(s, cudaEventRecord(s->evStateFree, s->stream));
                BH_CUDA(s, cudaStreamWaitEvent(up, s->evStateFree, 0));
                s->stagingSharedUse = false;
            }
        }
        for (int k = 0; k < 7; ++k) {
            const int i = order[k];
            BH_CUDA(s, cudaMemcpyAsync(stg + i * n, src[i], sizeof(float) * n, cudaMemcpyHostToDevice, up));
            dev[i] = stg + i * n;
            if (k == 3) {
                // the position copies gate the next step: a pending vertex read-back (third stream) starts behind them
                BH_CUDA(s, cudaEventRecord(s->evPosArrived, up));
                rc = flushCopy(s, s->evPosArrived);
                if (rc) return rc;
            }
        }
        if (async) BH_CUDA(s, cudaStreamWaitEvent(s->stream, s->evPosArrived, 0));
    } else {
        for (int i = 0; i < 7; ++i) dev[i] = src[i];
    }
    if (async) {  // the velocities of a previous asynchronous upload are part of the state that is read next
        rc = settleVel(s);
        if (rc) return rc;
    }
    // Placement: body i goes to the slot the body with the same host number has in the state being replaced (tree order
    // of the last step), so that a host that sends its bodies every step keeps the tree stages' locality.
    const int *slotOf = nullptr;
    if (s->placed) {
        rc = ensureSlotOf(s);
        if (rc) return rc;
        slotOf = s->slotOf;  // (and it describes the new state as well: body i goes where body i was)
    }
    if (async) {  // pack_vel overwrites the state: everything enqueued on the simulation's stream so far reads the old one
        BH_CUDA(s, cudaEventRecord(s->evStateFree, s->stream));
        BH_CUDA(s, cudaStreamWaitEvent(s->upStream, s->evStateFree, 0));
    }
    rc = resetState(s);
    if (rc) return rc;
    bh::pack_pos_kernel<<<grid, 256, 0, s->stream>>>(dev[0], dev[1], dev[2], dev[6], s->body4[0], slotOf, s->perm, s->n);
    BH_CUDA(s, cudaGetLastError());
    bh::pack_vel_kernel<<<grid, 256, 0, async ? s->upStream : s->stream>>>(dev[3], dev[4], dev[5], s->velacc[0], slotOf, s->n);
    BH_CUDA(s, cudaGetLastError());
    s->placed = true;
    s->slotOfValid = slotOf != nullptr;  // placed by slotOf[]: it describes the new state as well
    if (async) {
        BH_CUDA(s, cudaEventRecord(s->evPosPacked, s->stream));
        BH_CUDA(s, cudaStreamWaitEvent(s->upStream, s->evPosPacked, 0));  // the staging buffer is free when BOTH packs are done
        BH_CUDA(s, cudaEventRecord(s->evVelReady, s->upStream));
        BH_CUDA(s, cudaEventRecord(s->evPosPacked, s->upStream));
        s->stagingBusy = true;
        s->velPending = true;
        return BH_OK;
    }
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    return BH_OK;
}

int bh_upload(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
              const float *vz, const float *mass) {
    BH_ENTER(sim);
    const float *src[7] = {x, y, z, vx, vy, vz, mass};
    return uploadImpl(s, src, cudaMemcpyHostToDevice, false);
}

int bh_upload_async(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                    const float *vz, const float *mass) {
    BH_ENTER(sim);
    const float *src[7] = {x, y, z, vx, vy, vz, mass};
    return uploadImpl(s, src, cudaMemcpyHostToDevice, true);
}

int bh_upload_device(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                     const float *vz, const float *mass) {
    BH_ENTER(sim);
    const float *src[7] = {x, y, z, vx, vy, vz, mass};
    return uploadImpl(s, src, cudaMemcpyDeviceToDevice, false);
}

int bh_bounding_box(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_BBOX); }
int bh_build_tree(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_BUILD); }
int bh_summarize(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_SUMMARIZE); }
int bh_sort(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_SORT); }
int bh_calculate_force(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_FORCE); }
int bh_integrate(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_INTEGRATE); }

int bh_stage_async(bh_sim *sim, int32_t stage) { BH_ENTER(sim); return launchStage(s, stage, false); }

int bh_step_async(bh_sim *sim, int32_t nsteps) {
    BH_ENTER(sim);
    if (nsteps < 0) return fail(s, BH_ERR_ARG, "nsteps < 0");
    return stepAsync(s, nsteps);
}

int bh_check(bh_sim *sim) { BH_ENTER(sim); return finish(s); }

int bh_step(bh_sim *sim, int32_t nsteps) {
    BH_ENTER(sim);
    if (nsteps < 0) return fail(s, BH_ERR_ARG, "nsteps < 0");
    // chunks of kProfSteps so that per-stage events never run out
    while (nsteps > 0) {
        const int chunk = s->profiling ? std::min<int>(nsteps, kProfSteps) : nsteps;
        int rc = stepAsync(s, chunk);
        if (rc) return rc;
        if (s->profiling || chunk == nsteps) {
            rc = finish(s);
            if (rc) return rc;
        }
        nsteps -= chunk;
    }
    return BH_OK;
}

int bh_calculate_force_slice(bh_sim *sim, int32_t first, int32_t count) {
    BH_ENTER(sim);
    if (!validSlice(s, first, count)) return fail(s, BH_ERR_ARG, "bad slice [%d, %d): first must be a multiple of vote_width", first, first + count);
    if (count == 0) return BH_OK;
    if ((s->stagesRun & (1u << BH_STAGE_SORT)) == 0) return fail(s, BH_ERR_ARG, "calculate_force called before the tree stages have run since the upload");
    launchWalk(s, first, count, false);
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[BH_STAGE_FORCE] += 1;
    return BH_OK;
}

int bh_calculate_force_slice_p2p(bh_sim *sim, int32_t first, int32_t count) {
    BH_ENTER(sim);
    if (!s->p2p) return fail(s, BH_ERR_ARG, "bh_ipc_set_peers has not been called");
    if (!validSlice(s, first, count)) return fail(s, BH_ERR_ARG, "bad slice [%d, %d): first must be a multiple of vote_width", first, first + count);
    if (count == 0) return BH_OK;
    if ((s->stagesRun & (1u << BH_STAGE_SORT)) == 0) return fail(s, BH_ERR_ARG, "calculate_force called before the tree stages have run since the upload");
    launchWalk(s, first, count, true);
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[BH_STAGE_FORCE] += 1;
    return BH_OK;
}

int bh_peer_barrier(bh_sim *sim) {
    BH_ENTER(sim);
    if (!s->p2p) return fail(s, BH_ERR_ARG, "bh_ipc_set_peers has not been called");
    return launchBarrier(s);
}

int bh_apply_acceleration(bh_sim *sim) {
    BH_ENTER_SETTLED(sim);
    const unsigned stride = s->p2p ? (unsigned)accStride(s) : 0u;
    bh::apply_acc_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->acc, stride, s->permValid ? s->perm : nullptr, s->velacc[s->cur],
                                                                   s->sc, s->n, s->dt);
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[BH_STAGE_FORCE]++;
    return BH_OK;
}

int bh_finish_async(bh_sim *sim) {
    BH_ENTER_SETTLED(sim);
    int rc = launchFinish(s, true);
    if (rc == BH_OK) s->stageLaunches[BH_STAGE_INTEGRATE]++;
    return rc;
}

void *bh_acc_sorted_device_ptr(bh_sim *sim) { return sim ? S(sim)->acc : nullptr; }

int bh_set_slice(bh_sim *sim, int32_t first, int32_t count) {
    BH_ENTER(sim);
    if (!validSlice(s, first, count)) return fail(s, BH_ERR_ARG, "bad slice [%d, %d): first must be a multiple of vote_width", first, first + count);
    if (!s->p2p && (first != 0 || count != s->n))
        return fail(s, BH_ERR_ARG, "a slice smaller than the universe needs peer memory (bh_ipc_set_peers): the step's all-gather runs over it");
    dropGraphs(s);
    s->sliceFirst = first;
    s->sliceCount = count;
    return BH_OK;
}

int bh_ipc_export(bh_sim *sim, void *handle64) {
    BH_ENTER(sim);
    if (!handle64) return fail(s, BH_ERR_ARG, "handle64 is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    BH_CUDA(s, cudaIpcGetMemHandle(&h, s->accAlloc));
    memcpy(handle64, &h, sizeof h);
    return BH_OK;
}

int bh_ipc_clear_peers(bh_sim *sim) {
    BH_ENTER(sim);
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    dropGraphs(s);
    closePeers(s);
    s->sliceFirst = 0;
    s->sliceCount = s->n;
    return BH_OK;
}

int bh_ipc_set_peers(bh_sim *sim, int32_t nranks, int32_t my_rank, const void *handles) {
    BH_ENTER(sim);
    if (nranks < 1 || nranks > bh::kMaxPeers || my_rank < 0 || my_rank >= nranks || !handles)
        return fail(s, BH_ERR_ARG, "bad peer set (at most %d ranks)", bh::kMaxPeers);
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    dropGraphs(s);
    closePeers(s);
    for (int r = 0; r < nranks; ++r) {
        if (r == my_rank) { s->peerAlloc[r] = s->accAlloc; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + 64 * (size_t)r, sizeof h);
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            closePeers(s);  // release what was opened: the caller falls back to its own all-gather
            return fail(s, BH_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
        }
        s->peerAlloc[r] = static_cast<char *>(p);
    }
    s->nranks = nranks;
    s->rank = my_rank;
    s->p2p = nranks > 1;
    return BH_OK;
}

int64_t bh_buffer_length(bh_sim *sim, int32_t which) {
    if (!sim) return BH_ERR_ARG;
    Sim *s = S(sim);
    const int64_t m1 = (int64_t)s->m + 1;
    switch (which) {
    case BH_STEP: case BH_BLOCK_COUNT: case BH_RADIUS: case BH_MAX_DEPTH: case BH_BOTTOM: case BH_ERROR: return 1;
    case BH_CHILD: return 8 * m1;
    default: return (which >= 0 && which < BH_NUM_BUFFERS) ? m1 : (int64_t)BH_ERR_ARG;
    }
}

int bh_read(bh_sim *sim, int32_t which, void *dst, int64_t count) {
    BH_ENTER_SETTLED(sim);
    const int64_t len = bh_buffer_length(sim, which);
    if (len < 0) return fail(s, BH_ERR_ARG, "unknown buffer %d", which);
    if (!dst || count < 0 || count > len) return fail(s, BH_ERR_ARG, "bad destination/count for buffer %d", which);
    if (count == 0) return BH_OK;
    if (len == 1) {
        BH_CUDA(s, cudaMemcpyAsync(s->hostSc, s->sc, sizeof(bh::Scalars), cudaMemcpyDeviceToHost, s->stream));
        BH_CUDA(s, cudaStreamSynchronize(s->stream));
        const bh::Scalars &h = *s->hostSc;
        switch (which) {
        case BH_STEP: *static_cast<int32_t *>(dst) = h.step; break;
        case BH_BLOCK_COUNT: *static_cast<int32_t *>(dst) = h.blockCount; break;
        case BH_RADIUS: *static_cast<float *>(dst) = h.radius; break;
        case BH_MAX_DEPTH: *static_cast<int32_t *>(dst) = h.maxDepth; break;
        case BH_BOTTOM: *static_cast<int32_t *>(dst) = h.bottom; break;
        default: *static_cast<int32_t *>(dst) = h.error; break;
        }
        return BH_OK;
    }
    int rc = ensureStaging(s, 4 * (size_t)count);
    if (rc) return rc;
    const unsigned grid = (unsigned)((count + 255) / 256);
    const unsigned gridBodies = (unsigned)((std::max<int64_t>(count, s->n) + 255) / 256);  // bodies scatter to the host's numbering
    float *fout = static_cast<float *>(s->staging);
    int *iout = static_cast<int *>(s->staging);
    const float4 *body = s->body4[s->cur], *va = s->velacc[s->cur];
    switch (which) {
    case BH_POS_X: case BH_POS_Y: case BH_POS_Z:
        bh::export_float_kernel<<<gridBodies, 256, 0, s->stream>>>(body, va, s->cell4, which - BH_POS_X, fout, s->n, count);
        break;
    case BH_VEL_X: case BH_VEL_Y: case BH_VEL_Z:
        bh::export_float_kernel<<<gridBodies, 256, 0, s->stream>>>(body, va, s->cell4, 3 + which - BH_VEL_X, fout, s->n, count);
        break;
    case BH_ACC_X: case BH_ACC_Y: case BH_ACC_Z:
        bh::export_float_kernel<<<gridBodies, 256, 0, s->stream>>>(body, va, s->cell4, 6 + which - BH_ACC_X, fout, s->n, count);
        break;
    case BH_MASS:
        bh::export_float_kernel<<<gridBodies, 256, 0, s->stream>>>(body, va, s->cell4, 9, fout, s->n, count);
        break;
    case BH_BODY_COUNT:
        bh::export_shifted_kernel<<<grid, 256, 0, s->stream>>>(s->count, s->n, bh::kCountMask, iout, count);
        break;
    case BH_START:
        bh::export_shifted_kernel<<<grid, 256, 0, s->stream>>>(s->start, s->n, -1, iout, count);
        break;
    case BH_CHILD:
        bh::export_child_kernel<<<grid, 256, 0, s->stream>>>(s->child, s->velacc[s->treePhase], s->sc, s->n, iout, count);
        break;
    case BH_SORTED:
        if (!s->havePerm) BH_CUDA(s, cudaMemsetAsync(iout, 0, 4 * (size_t)count, s->stream));  // GPUBH:178: zeros until the first sort
        else bh::export_sorted_kernel<<<grid, 256, 0, s->stream>>>(s->perm, va, s->permValid ? 1 : 0, s->n, iout, count);
        break;
    default:
        return fail(s, BH_ERR_ARG, "unknown buffer %d", which);
    }
    BH_CUDA(s, cudaGetLastError());
    BH_CUDA(s, cudaMemcpyAsync(dst, s->staging, 4 * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    return BH_OK;
}

int bh_copy_vertices_device(bh_sim *sim, void *pos4_device, void *vel4_device) {
    BH_ENTER_SETTLED(sim);
    if (!pos4_device && !vel4_device) return BH_OK;
    int rc = exportVertices(s, static_cast<float4 *>(pos4_device), static_cast<float4 *>(vel4_device));
    if (rc) return rc;
    return BH_OK;
}

int bh_copy_vertices_async(bh_sim *sim, float *pos4, float *vel4) {
    BH_ENTER_SETTLED(sim);
    if (!pos4 && !vel4) return BH_OK;
    const size_t bytes = sizeof(float4) * (size_t)s->n;
    if (!s->vtxStaging) {
        BH_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->vtxStaging), 2 * bytes));
        BH_CUDA(s, cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
        BH_CUDA(s, cudaEventCreateWithFlags(&s->evExported, cudaEventDisableTiming));
        BH_CUDA(s, cudaEventCreateWithFlags(&s->evCopied, cudaEventDisableTiming));
    }
    int rc = flushCopy(s, nullptr);  // an earlier export that nobody has waited for
    if (rc) return rc;
    if (s->copyPending) BH_CUDA(s, cudaStreamWaitEvent(s->stream, s->evCopied, 0));  // the staging buffer is still being read
    float4 *dp = s->vtxStaging, *dv = dp + s->n;
    rc = exportVertices(s, pos4 ? dp : nullptr, vel4 ? dv : nullptr);
    if (rc) return rc;
    BH_CUDA(s, cudaEventRecord(s->evExported, s->stream));
    s->deferredPos = pos4;
    s->deferredVel = vel4;
    s->copyDeferred = true;
    return BH_OK;
}

int bh_wait_copies(bh_sim *sim) {
    BH_ENTER(sim);
    int rc = flushCopy(s, nullptr);
    if (rc) return rc;
    if (s->copyPending) {
        BH_CUDA(s, cudaEventSynchronize(s->evCopied));
        s->copyPending = false;
    }
    return BH_OK;
}

int bh_copy_vertices(bh_sim *sim, float *pos4, float *vel4) {
    int rc = bh_copy_vertices_async(sim, pos4, vel4);
    return rc ? rc : bh_wait_copies(sim);
}

int bh_stats(bh_sim *sim, bh_stats_t *out) {
    BH_ENTER_SETTLED(sim);
    if (!out) return fail(s, BH_ERR_ARG, "out is NULL");
    BH_CUDA(s, cudaMemcpyAsync(s->hostSc, s->sc, sizeof(bh::Scalars), cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    memset(out, 0, sizeof *out);
    out->nbodies = s->n;
    out->number_of_nodes = s->m;
    out->cells_used = s->m - s->hostSc->bottom + 1;
    out->max_depth = s->hostSc->maxDepth;
    out->step = s->hostSc->step;
    out->error = s->hostSc->error;
    out->steps_timed = s->stepsTimed;
    // events are recorded in launch order: slots 0-4 = bbox..force, 5 = barrier, 6 = finish
    for (int i = 0; i < BH_NUM_STAGES; ++i) {
        out->stage_ms[i] = s->stageMs[i == BH_STAGE_INTEGRATE ? kNumTimers - 1 : i];
        out->stage_launches[i] = s->stageLaunches[i];
    }
    out->barrier_ms = s->stageMs[BH_STAGE_INTEGRATE];
    out->interactions = (int64_t)s->hostSc->interactions;
    out->opens = (int64_t)s->hostSc->opens;
    out->deep_walk = (s->vote != 16 || s->forceDeep) ? 1 : 0;
    out->walk_spills = s->hostSc->walkSpills;
    return BH_OK;
}

int bh_reset_stats(bh_sim *sim) {
    BH_ENTER(sim);
    for (auto &v : s->stageMs) v = 0.0;
    for (auto &v : s->stageLaunches) v = 0;
    s->stepsTimed = 0;
    return BH_OK;
}

int32_t bh_number_of_bodies(bh_sim *sim) { return sim ? S(sim)->n : BH_ERR_ARG; }

int bh_generate_universe(bh_sim *sim, int32_t kind, uint64_t seed, float p0, float p1, float p2) {
    BH_ENTER_SETTLED(sim);
    if (kind < 0 || kind > 2) return fail(s, BH_ERR_ARG, "unknown universe kind %d", kind);
    int rc = resetState(s);
    if (rc) return rc;
    bh::generate_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->body4[0], s->velacc[0], s->perm, s->n, kind, seed, p0, p1, p2);
    BH_CUDA(s, cudaGetLastError());
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    s->placed = true;
    s->slotOfValid = false;
    return BH_OK;
}

int bh_diagnostics(bh_sim *sim, int32_t with_potential, bh_diag_t *out) {
    BH_ENTER_SETTLED(sim);
    if (!out) return fail(s, BH_ERR_ARG, "out is NULL");
    int rc = ensureStaging(s, 8 * sizeof(double));
    if (rc) return rc;
    double *d = static_cast<double *>(s->staging);
    BH_CUDA(s, cudaMemsetAsync(d, 0, 8 * sizeof(double), s->stream));
    bh::kinetic_kernel<<<std::min((s->n + 255) / 256, s->numSMs * 8), 256, 0, s->stream>>>(s->body4[s->cur], s->velacc[s->cur], d, s->n);
    if (with_potential == 1) {
        bh::potential_kernel<<<(s->n + bh::kPotTile - 1) / bh::kPotTile, bh::kPotTile, 0, s->stream>>>(s->body4[s->cur], d, s->n, s->eps);
    } else if (with_potential == 2) {
        // the tree's potential: tree stages + the walk's potential variant on the current positions; the bodies, the step
        // counter and the accelerations of the state are left as they are (only the tree buffers and the scratch
        // acceleration buffer are overwritten, as by the next step)
        if (s->vote != 16) return fail(s, BH_ERR_ARG, "the tree potential needs vote_width 16");
        const bool counting = s->counting;
        s->counting = false;
        int rcs = BH_OK;
        for (int st = BH_STAGE_BBOX; st <= BH_STAGE_SORT && rcs == BH_OK; ++st) {
            rcs = launchStage(s, st, true);
            if (st == BH_STAGE_BBOX) bh::adjust_step_kernel<<<1, 1, 0, s->stream>>>(s->sc, -1);  // boundingbox.cl:195 counted a step
        }
        s->counting = counting;
        if (rcs) return rcs;
        launchWalk(s, 0, s->n, false, true);
        bh::tree_potential_kernel<<<std::min((s->n + 255) / 256, s->numSMs * 8), 256, 0, s->stream>>>(
            s->body4[s->cur], s->permValid ? s->perm : nullptr, s->acc, s->p2p ? (unsigned)accStride(s) : 0u, s->sc, d, s->n, s->eps);
    }
    BH_CUDA(s, cudaGetLastError());
    double h[8];
    BH_CUDA(s, cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    out->ekin = h[0]; out->px = h[1]; out->py = h[2]; out->pz = h[3]; out->mass = h[4];
    out->epot = with_potential ? h[5] : 0.0;
    return BH_OK;
}

// ---- .universe files (Java ObjectOutputStream layout, UniverseSerializer.java:25-34) ----
namespace {
const unsigned char kMagic[6] = {0xAC, 0xED, 0x00, 0x05, 0x77, 0x04};  // stream header + TC_BLOCKDATA(4) = writeInt
const unsigned char kClassDesc[] = {0x72, 0x00, 0x02, '[', 'F', 0x0B, 0x9C, 0x81, 0x89, 0x22, 0xE0, 0x0C, 0x42, 0x02, 0x00, 0x00, 0x78, 0x70};
const unsigned char kClassRef[] = {0x71, 0x00, 0x7E, 0x00, 0x00};

struct UniverseFile {
    int32_t n = 0;
    std::string error;
    std::vector<float> arrays[7];
};
uint32_t be32(const unsigned char *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
void putBe32(unsigned char *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

// expect > 0: the body count the caller needs (checked before anything is allocated); 0 = any
bool readUniverseBody(FILE *f, bool headerOnly, int32_t expect, UniverseFile &u) {
    unsigned char hdr[10];
    if (fread(hdr, 1, 10, f) != 10 || memcmp(hdr, kMagic, 6) != 0) { u.error = "not a .universe file (Java stream with a leading writeInt expected)"; return false; }
    u.n = (int32_t)be32(hdr + 6);
    if (u.n <= 0) { u.error = "invalid body count in the file header"; return false; }
    if (headerOnly) return true;
    if (expect > 0 && u.n != expect) {  // SerializedUniverseGenerator.java:41-42 IllegalStateException
        char msg[160];
        snprintf(msg, sizeof msg, "invalid amount of bodies for serialized universe (%d in the file, %d in the simulation)", u.n, expect);
        u.error = msg;
        return false;
    }
    std::vector<unsigned char> buf(4 * (size_t)u.n);
    for (int a = 0; a < 7; ++a) {
        unsigned char tag[32];
        if (fread(tag, 1, 2, f) != 2 || tag[0] != 0x75) { u.error = "expected TC_ARRAY"; return false; }
        const bool full = tag[1] == 0x72;
        const size_t rest = (full ? sizeof kClassDesc : sizeof kClassRef) - 1;
        if (fread(tag + 2, 1, rest, f) != rest || memcmp(tag + 1, full ? kClassDesc : kClassRef, rest + 1) != 0) {
            u.error = "unexpected array class (float[] expected)";
            return false;
        }
        unsigned char len[4];
        if (fread(len, 1, 4, f) != 4 || (int32_t)be32(len) != u.n) { u.error = "array length differs from nbodies"; return false; }
        if (fread(buf.data(), 1, buf.size(), f) != buf.size()) { u.error = "short read"; return false; }
        u.arrays[a].resize((size_t)u.n);
        for (int32_t i = 0; i < u.n; ++i) {
            const uint32_t v = be32(buf.data() + 4 * (size_t)i);
            memcpy(&u.arrays[a][i], &v, 4);
        }
    }
    return true;
}

bool readUniverse(const char *path, bool headerOnly, int32_t expect, UniverseFile &u) {
    FILE *f = fopen(path, "rb");
    if (!f) { u.error = std::string("cannot open ") + path; return false; }
    bool ok = false;
    try {
        ok = readUniverseBody(f, headerOnly, expect, u);
    } catch (const std::exception &) {  // bad_alloc / length_error must not cross the C boundary
        u.error = "out of host memory while reading the universe file";
    }
    fclose(f);
    return ok;
}
}  // namespace

int bh_universe_file_bodies(const char *path, int32_t *nbodies) {
    if (!path || !nbodies) return BH_ERR_ARG;
    UniverseFile u;
    if (!readUniverse(path, true, 0, u)) return fail(nullptr, BH_ERR_ARG, "%s", u.error.c_str());
    *nbodies = u.n;
    return BH_OK;
}

int bh_read_universe_file(const char *path, int32_t capacity, float *x, float *y, float *z, float *vx, float *vy, float *vz, float *mass) {
    float *dst[7] = {x, y, z, vx, vy, vz, mass};
    if (!path) return fail(nullptr, BH_ERR_ARG, "path is NULL");
    for (float *d : dst)
        if (!d) return fail(nullptr, BH_ERR_ARG, "NULL output array");
    UniverseFile u;
    if (!readUniverse(path, true, 0, u)) return fail(nullptr, BH_ERR_ARG, "%s", u.error.c_str());
    if (u.n > capacity) return fail(nullptr, BH_ERR_ARG, "%d bodies in the file, room for %d", u.n, capacity);
    if (!readUniverse(path, false, 0, u)) return fail(nullptr, BH_ERR_ARG, "%s", u.error.c_str());
    for (int a = 0; a < 7; ++a) memcpy(dst[a], u.arrays[a].data(), sizeof(float) * (size_t)u.n);
    return BH_OK;
}

int bh_upload_universe_file(bh_sim *sim, const char *path) {
    BH_ENTER(sim);
    if (!path) return fail(s, BH_ERR_ARG, "path is NULL");
    UniverseFile u;
    if (!readUniverse(path, false, s->n, u)) return fail(s, BH_ERR_ARG, "%s", u.error.c_str());
    const float *src[7];
    for (int a = 0; a < 7; ++a) src[a] = u.arrays[a].data();
    return uploadImpl(s, src, cudaMemcpyHostToDevice, false);
}

int bh_write_universe_file(bh_sim *sim, const char *path) {
    BH_ENTER_SETTLED(sim);
    if (!path) return fail(s, BH_ERR_ARG, "path is NULL");
    const size_t n = s->n;
    int rc = ensureStaging(s, sizeof(float) * 7 * n);
    if (rc) return rc;
    bh::export_universe_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->body4[s->cur], s->velacc[s->cur], static_cast<float *>(s->staging), s->n);
    BH_CUDA(s, cudaGetLastError());
    try {
        std::vector<float> host(7 * n);
        BH_CUDA(s, cudaMemcpyAsync(host.data(), s->staging, sizeof(float) * 7 * n, cudaMemcpyDeviceToHost, s->stream));
        BH_CUDA(s, cudaStreamSynchronize(s->stream));
        FILE *f = fopen(path, "wb");
        if (!f) return fail(s, BH_ERR_ARG, "cannot create %s", path);
        std::vector<unsigned char> buf(4 * n);
        unsigned char word[4];
        bool ok = fwrite(kMagic, 1, 6, f) == 6;
        putBe32(word, (uint32_t)s->n);
        ok = ok && fwrite(word, 1, 4, f) == 4;
        for (int a = 0; a < 7 && ok; ++a) {  // UniverseSerializer.java:27-33: x, y, z, vx, vy, vz, mass
            const unsigned char tcArray = 0x75;
            ok = fwrite(&tcArray, 1, 1, f) == 1;
            if (a == 0) ok = ok && fwrite(kClassDesc, 1, sizeof kClassDesc, f) == sizeof kClassDesc;
            else ok = ok && fwrite(kClassRef, 1, sizeof kClassRef, f) == sizeof kClassRef;
            ok = ok && fwrite(word, 1, 4, f) == 4;
            for (size_t i = 0; i < n; ++i) {
                uint32_t v;
                memcpy(&v, &host[a * n + i], 4);
                putBe32(buf.data() + 4 * i, v);
            }
            ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        }
        ok = (fclose(f) == 0) && ok;
        if (!ok) return fail(s, BH_ERR_ARG, "short write to %s", path);
    } catch (const std::exception &) {
        return fail(s, BH_ERR_ALLOC, "out of host memory while writing the universe file");
    }
    return BH_OK;
}

static int measurePeak(int32_t device, int packed, double *tflops) {
    if (!tflops) return BH_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, BH_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BH_ERR_CUDA;
    float *out = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&out), 4) != cudaSuccess) return BH_ERR_ALLOC;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        if (packed) bh::fp32x2_peak_kernel<<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f);
        else bh::fp32_peak_kernel<<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = (packed ? 4.0 : 2.0) * 64.0 * iters * 256.0 * grid;
        if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *tflops = best;
    return best > 0.0 ? BH_OK : BH_ERR_CUDA;
}

int bh_measure_fp32_peak(int32_t device, double *tflops) { return measurePeak(device, 0, tflops); }
int bh_measure_fp32x2_rate(int32_t device, double *tflops) { return measurePeak(device, 1, tflops); }

}  // extern "C"
