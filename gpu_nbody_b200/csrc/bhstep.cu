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

#include "bhstep.h"

namespace {

constexpr int kProfSteps = 64;  // steps per bh_step call that get per-stage events
constexpr size_t kAccPad = 2048; // slack of the sorted-order acceleration buffer: equal, 32-aligned slices for up to 64 ranks

thread_local std::string g_createError;

struct Sim {
    int device = 0;
    int n = 0, m = 0, nc = 0;  // bodies, NUMBER_OF_NODES, cell slots (m - n + 1)
    float theta = 0.5f, thetaMacro = 0.25f, eps = 0.0025f, dt = 0.025f;
    int vote = 16;
    int numSMs = 0;
    cudaStream_t stream = nullptr, ownStream = nullptr;
    // device state
    float4 *node4 = nullptr, *velacc = nullptr, *octet = nullptr, *accSorted = nullptr;
    int *child = nullptr, *start = nullptr, *count = nullptr, *sorted = nullptr, *meta = nullptr, *oidx = nullptr, *parent = nullptr, *arrived = nullptr;
    float *partials = nullptr;
    bh::Scalars *sc = nullptr;
    bh::Scalars *hostSc = nullptr;  // pinned mirror
    void *staging = nullptr;
    size_t stagingBytes = 0;
    // launch geometry
    int bboxGrid = 1, buildGrid = 1, summGrid = 1, sortGrid = 1;
    // options
    bool profiling = false, counting = false;
    int insertionOrder = 1;
    bool haveSorted = false;
    // profiling
    cudaEvent_t ev[kProfSteps][BH_NUM_STAGES + 1] = {};
    bool evCreated = false;
    int evSteps = 0;
    double stageMs[BH_NUM_STAGES] = {};
    int64_t stageLaunches[BH_NUM_STAGES] = {};
    int64_t stepsTimed = 0;
    // peer memory (multi-GPU): every rank's accSorted mapped through CUDA IPC; two phases (ping-pong per step)
    int nranks = 1, rank = 0, accPhase = 0;
    bool p2p = false;
    float4 *peerAcc[bh::kMaxPeers] = {};
    // CUDA graph of one step
    bool useGraph = true;
    cudaGraphExec_t graphExec = nullptr;
    cudaStream_t graphStream = nullptr;
    int graphInsertion = -1;
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

int ensureStaging(Sim *s, size_t bytes) {
    if (bytes <= s->stagingBytes) return BH_OK;
    if (s->staging) cudaFree(s->staging);
    s->staging = nullptr;
    s->stagingBytes = 0;
    BH_CUDA(s, cudaMalloc(&s->staging, bytes));
    s->stagingBytes = bytes;
    return BH_OK;
}

int resetState(Sim *s) {
    // GPUBH:155-179: everything zero except step = -1, maxDepth = 1
    bh::Scalars init;
    memset(&init, 0, sizeof init);
    init.step = -1;
    init.maxDepth = 1;
    *s->hostSc = init;
    BH_CUDA(s, cudaMemcpyAsync(s->sc, s->hostSc, sizeof init, cudaMemcpyHostToDevice, s->stream));
    BH_CUDA(s, cudaMemsetAsync(s->child, 0, sizeof(int) * 8 * (size_t)s->nc, s->stream));
    BH_CUDA(s, cudaMemsetAsync(s->start, 0, sizeof(int) * (size_t)s->nc, s->stream));
    BH_CUDA(s, cudaMemsetAsync(s->count, 0, sizeof(int) * (size_t)s->nc, s->stream));
    BH_CUDA(s, cudaMemsetAsync(s->node4 + s->n, 0, sizeof(float4) * (size_t)s->nc, s->stream));
    s->haveSorted = false;
    return BH_OK;
}

inline size_t accStride(const Sim *s) { return (size_t)s->n + kAccPad; }
inline float4 *accPhase(const Sim *s) { return s->accSorted + (size_t)s->accPhase * accStride(s); }

// force walk for sorted slots [first, first+cnt): fused velocity correction (slice = false) or
// sorted-order acceleration output (slice = true)
void launchForce(Sim *s, int first, int cnt, bool slice, bool counting, bool peers = false) {
    bh::PeerBuffers dst;
    dst.count = 1;
    dst.buf[0] = accPhase(s);
    if (peers)
        for (int r = 0; r < s->nranks; ++r)
            if (r != s->rank) dst.buf[dst.count++] = s->peerAcc[r] + (size_t)s->accPhase * accStride(s);
#define BH_FORCE_ARGS2 s->node4, s->octet, s->oidx, s->meta, s->sorted, s->velacc, dst, s->sc, s->n, s->m, first, cnt, s->thetaMacro, s->eps, s->dt
#define BH_FORCE_DISPATCH(KERNEL, THREADS, BODIES, ARGS)                                                     \
    do {                                                                                                     \
        const int grid = (cnt + (BODIES) - 1) / (BODIES);                                                     \
        if (s->vote == 16) {                                                                                 \
            if (slice) KERNEL<16, true, false><<<grid, THREADS, 0, s->stream>>>(ARGS);                       \
            else if (counting) KERNEL<16, false, true><<<grid, THREADS, 0, s->stream>>>(ARGS);               \
            else KERNEL<16, false, false><<<grid, THREADS, 0, s->stream>>>(ARGS);                            \
        } else {                                                                                             \
            if (slice) KERNEL<32, true, false><<<grid, THREADS, 0, s->stream>>>(ARGS);                       \
            else if (counting) KERNEL<32, false, true><<<grid, THREADS, 0, s->stream>>>(ARGS);               \
            else KERNEL<32, false, false><<<grid, THREADS, 0, s->stream>>>(ARGS);                            \
        }                                                                                                    \
    } while (0)
    BH_FORCE_DISPATCH(bh::force2_kernel, bh::kForce2Threads, bh::kForce2Bodies, BH_FORCE_ARGS2);
#undef BH_FORCE_DISPATCH
#undef BH_FORCE_ARGS2
}

int launchStage(Sim *s, int stage) {
    const int n = s->n, m = s->m;
    switch (stage) {
    case BH_STAGE_BBOX:
        bh::bbox_kernel<<<s->bboxGrid, bh::kBboxThreads, 0, s->stream>>>(s->node4, s->child, s->start, s->count, s->arrived,
                                                                         s->partials, s->sc, n, m);
        break;
    case BH_STAGE_BUILD:
        bh::build_kernel<<<s->buildGrid, bh::kBuildThreads, 0, s->stream>>>(
            s->node4, s->child, s->start, s->count, s->parent, s->arrived,
            (s->insertionOrder == 1 && s->haveSorted) ? s->sorted : nullptr, s->sc, n, m);
        break;
    case BH_STAGE_SUMMARIZE:
        bh::summarize_kernel<<<s->summGrid, bh::kSummThreads, 0, s->stream>>>(s->node4, s->child, s->octet, s->oidx, s->meta, s->count,
                                                                              s->parent, s->arrived, s->sc, n, m);
        break;
    case BH_STAGE_SORT:
        bh::sort_kernel<<<s->sortGrid, bh::kSortThreads, 0, s->stream>>>(s->child, s->count, s->start, s->sorted, s->sc, n, m);
        s->haveSorted = true;
        break;
    case BH_STAGE_FORCE: {
        if (s->counting) BH_CUDA(s, cudaMemsetAsync(&s->sc->interactions, 0, 2 * sizeof(unsigned long long), s->stream));
        launchForce(s, 0, n, false, s->counting);
        break;
    }
    case BH_STAGE_INTEGRATE:
        bh::integrate_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->node4, s->velacc, s->sc, n, s->dt);
        break;
    default:
        return fail(s, BH_ERR_ARG, "unknown stage %d", stage);
    }
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[stage]++;
    return BH_OK;
}

// sync, pull the scalars, report the device error buffer
int finish(Sim *s) {
    BH_CUDA(s, cudaMemcpyAsync(s->hostSc, s->sc, sizeof(bh::Scalars), cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    if (s->profiling && s->evSteps > 0) {
        for (int i = 0; i < s->evSteps; ++i)
            for (int st = 0; st < BH_NUM_STAGES; ++st) {
                float ms = 0.0f;
                if (cudaEventElapsedTime(&ms, s->ev[i][st], s->ev[i][st + 1]) == cudaSuccess) s->stageMs[st] += ms;
            }
        s->stepsTimed += s->evSteps;
        s->evSteps = 0;
    }
    if (s->hostSc->error != 0) {
        fail(s, s->hostSc->error, "device error buffer = %d (%s)", s->hostSc->error,
             s->hostSc->error == 1 ? "cell pool exhausted or tree deeper than 64 levels" : "device-side wait exceeded its spin budget");
        return s->hostSc->error;
    }
    return BH_OK;
}

// One step = six launches with fixed arguments once a sorted order exists: captured once into a CUDA graph and
// replayed (one launch call per step instead of six; matters for small universes such as the reference's 32 768-body
// default, where a step is a few hundred microseconds).  Not used while per-stage events or counters are on.
int graphStep(Sim *s) {
    if (!s->graphExec || s->graphStream != s->stream || s->graphInsertion != s->insertionOrder) {
        if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();    // e.g. a stream that cannot be captured: plain launches from now on
            s->useGraph = false;
            return BH_ERR_ARG;     // tells stepAsync to launch this step the ordinary way
        }
        int rc = BH_OK;
        for (int st = 0; st < BH_NUM_STAGES && rc == BH_OK; ++st) rc = launchStage(s, st);
        const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc != BH_OK || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc != BH_OK ? rc : fail(s, BH_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
        }
        for (int st = 0; st < BH_NUM_STAGES; ++st) s->stageLaunches[st]--;  // the capture did not run anything
        const cudaError_t ei = cudaGraphInstantiate(&s->graphExec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { s->graphExec = nullptr; return fail(s, BH_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); }
        s->graphStream = s->stream;
        s->graphInsertion = s->insertionOrder;
    }
    BH_CUDA(s, cudaGraphLaunch(s->graphExec, s->stream));
    for (int st = 0; st < BH_NUM_STAGES; ++st) s->stageLaunches[st]++;
    return BH_OK;
}

int stepAsync(Sim *s, int nsteps) {
    for (int i = 0; i < nsteps; ++i) {
        // (the legacy default stream cannot be captured)
        if (s->useGraph && !s->profiling && !s->counting && s->haveSorted && s->stream != nullptr) {
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
        for (int st = 0; st < BH_NUM_STAGES; ++st) {
            if (prof) BH_CUDA(s, cudaEventRecord(s->ev[s->evSteps][st], s->stream));
            int rc = launchStage(s, st);
            if (rc) return rc;
        }
        if (prof) {
            BH_CUDA(s, cudaEventRecord(s->ev[s->evSteps][BH_NUM_STAGES], s->stream));
            s->evSteps++;
        }
    }
    return BH_OK;
}

int singleStage(Sim *s, int stage) {
    const bool prof = s->profiling && s->evSteps < kProfSteps;
    if (prof) {
        // single stages are timed through one-off events folded into the same accumulators
        cudaEvent_t a, b;
        BH_CUDA(s, cudaEventCreate(&a));
        BH_CUDA(s, cudaEventCreate(&b));
        BH_CUDA(s, cudaEventRecord(a, s->stream));
        int rc = launchStage(s, stage);
        if (rc) return rc;
        BH_CUDA(s, cudaEventRecord(b, s->stream));
        rc = finish(s);
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) s->stageMs[stage] += ms;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        return rc;
    }
    int rc = launchStage(s, stage);
    if (rc) return rc;
    return finish(s);
}

}  // namespace

extern "C" {

int32_t bh_abi_version(void) { return 1; }

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
    // child rows are addressed as 8*(cell-N) in size_t; node indices must fit int32
    if (m < 0) return fail(nullptr, BH_ERR_ARG, "nbodies too large for 32-bit node indices");
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
    if ((e = cudaStreamCreateWithFlags(&s->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return bail(BH_ERR_CUDA, "cudaStreamCreate", e);
    s->stream = s->ownStream;
    const size_t n = nbodies, nc = s->nc;
#define BH_ALLOC(ptr, bytes) \
    if ((e = cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes))) != cudaSuccess) return bail(BH_ERR_ALLOC, "cudaMalloc " #ptr, e)
    BH_ALLOC(s->node4, sizeof(float4) * ((size_t)m + 1));
    BH_ALLOC(s->velacc, sizeof(float4) * 2 * n);
    BH_ALLOC(s->octet, sizeof(float4) * 8 * nc);
    BH_ALLOC(s->accSorted, sizeof(float4) * 2 * (n + kAccPad));  // two phases, see bh_ipc_set_peers
    BH_ALLOC(s->child, sizeof(int) * 8 * nc);
    BH_ALLOC(s->start, sizeof(int) * nc);
    BH_ALLOC(s->count, sizeof(int) * nc);
    BH_ALLOC(s->meta, sizeof(int) * nc);
    BH_ALLOC(s->oidx, sizeof(int) * 8 * nc);
    BH_ALLOC(s->parent, sizeof(int) * nc);
    BH_ALLOC(s->arrived, sizeof(int) * nc);
    BH_ALLOC(s->sorted, sizeof(int) * n);
    BH_ALLOC(s->sc, sizeof(bh::Scalars));
#undef BH_ALLOC
    if ((e = cudaMallocHost(reinterpret_cast<void **>(&s->hostSc), sizeof(bh::Scalars))) != cudaSuccess) return bail(BH_ERR_ALLOC, "cudaMallocHost", e);
    // launch geometry: streaming kernels a few CTAs per SM; the two kernels that wait on
    // other threads (summarise, sort) exactly as many CTAs as are resident at once.
    int perSM = 1;
    s->bboxGrid = (int)std::min<size_t>((n + bh::kBboxThreads - 1) / bh::kBboxThreads, (size_t)s->numSMs * 4);
    if ((e = cudaMalloc(reinterpret_cast<void **>(&s->partials), sizeof(float) * 6 * s->bboxGrid)) != cudaSuccess) return bail(BH_ERR_ALLOC, "cudaMalloc partials", e);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, bh::build_kernel, bh::kBuildThreads, 0);
    s->buildGrid = (int)std::min<size_t>((n + bh::kBuildThreads - 1) / bh::kBuildThreads, (size_t)s->numSMs * std::max(perSM, 1));
    s->summGrid = (int)std::min<size_t>((nc + bh::kSummThreads - 1) / bh::kSummThreads, (size_t)s->numSMs * 64);  // never waits: any grid
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, bh::sort_kernel, bh::kSortThreads, 0);
    s->sortGrid = s->numSMs * std::max(perSM, 1);
    // tiny problems: do not launch more waiting threads than there can be cells
    const int cellBlocks = (int)((nc + bh::kSummThreads - 1) / bh::kSummThreads);
    s->sortGrid = std::max(1, std::min(s->sortGrid, cellBlocks));
    if ((e = cudaMemsetAsync(s->node4, 0, sizeof(float4) * ((size_t)m + 1), s->stream)) != cudaSuccess) return bail(BH_ERR_CUDA, "cudaMemset", e);
    cudaMemsetAsync(s->velacc, 0, sizeof(float4) * 2 * n, s->stream);
    cudaMemsetAsync(s->sorted, 0, sizeof(int) * n, s->stream);
    cudaMemsetAsync(s->accSorted, 0, sizeof(float4) * 2 * (n + kAccPad), s->stream);
    if (resetState(s) != BH_OK || cudaStreamSynchronize(s->stream) != cudaSuccess) {
        g_createError = s->lastError.empty() ? "initial reset failed" : s->lastError;
        bh_destroy(reinterpret_cast<bh_sim *>(s));
        return BH_ERR_CUDA;
    }
    *out = reinterpret_cast<bh_sim *>(s);
    return BH_OK;
}

void bh_destroy(bh_sim *sim) {
    if (!sim) return;
    Sim *s = S(sim);
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->graphExec) cudaGraphExecDestroy(s->graphExec);
    for (int r = 0; r < s->nranks; ++r)
        if (s->p2p && r != s->rank && s->peerAcc[r]) cudaIpcCloseMemHandle(s->peerAcc[r]);
    cudaFree(s->node4); cudaFree(s->velacc); cudaFree(s->octet); cudaFree(s->accSorted);
    cudaFree(s->child); cudaFree(s->start); cudaFree(s->count); cudaFree(s->sorted); cudaFree(s->meta); cudaFree(s->oidx); cudaFree(s->parent); cudaFree(s->arrived);
    cudaFree(s->partials); cudaFree(s->sc); cudaFree(s->staging);
    if (s->hostSc) cudaFreeHost(s->hostSc);
    if (s->evCreated)
        for (auto &row : s->ev)
            for (auto &e : row) cudaEventDestroy(e);
    if (s->ownStream) cudaStreamDestroy(s->ownStream);
    delete s;
}

#define BH_ENTER(sim)                                          \
    if (!(sim)) return BH_ERR_ARG;                             \
    Sim *s = S(sim);                                           \
    BH_CUDA(s, cudaSetDevice(s->device))

int bh_set_theta_macro(bh_sim *sim, float theta_macro) {
    BH_ENTER(sim);
    if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }  // kernel arguments change
    s->thetaMacro = theta_macro;
    return BH_OK;
}

int bh_set_stream(bh_sim *sim, void *cuda_stream) {
    BH_ENTER(sim);
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    s->stream = reinterpret_cast<cudaStream_t>(cuda_stream);  // NULL = CUDA's default stream, as everywhere in CUDA
    return BH_OK;
}

int bh_use_private_stream(bh_sim *sim) {
    BH_ENTER(sim);
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
    s->insertionOrder = mode;
    return BH_OK;
}

static int uploadImpl(Sim *s, const float *const src[7], cudaMemcpyKind kind) {
    for (int i = 0; i < 7; ++i)
        if (!src[i]) return fail(s, BH_ERR_ARG, "NULL input array %d", i);
    const size_t n = s->n;
    const float *dev[7];
    if (kind == cudaMemcpyHostToDevice) {
        int rc = ensureStaging(s, sizeof(float) * 7 * n);
        if (rc) return rc;
        float *stg = static_cast<float *>(s->staging);
        for (int i = 0; i < 7; ++i) {
            BH_CUDA(s, cudaMemcpyAsync(stg + i * n, src[i], sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
            dev[i] = stg + i * n;
        }
    } else {
        for (int i = 0; i < 7; ++i) dev[i] = src[i];
    }
    int rc = resetState(s);
    if (rc) return rc;
    bh::pack_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], s->node4,
                                                              s->velacc, s->sorted, s->n);
    BH_CUDA(s, cudaGetLastError());
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    return BH_OK;
}

int bh_upload(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
              const float *vz, const float *mass) {
    BH_ENTER(sim);
    const float *src[7] = {x, y, z, vx, vy, vz, mass};
    return uploadImpl(s, src, cudaMemcpyHostToDevice);
}

int bh_upload_device(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                     const float *vz, const float *mass) {
    BH_ENTER(sim);
    const float *src[7] = {x, y, z, vx, vy, vz, mass};
    return uploadImpl(s, src, cudaMemcpyDeviceToDevice);
}

int bh_bounding_box(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_BBOX); }
int bh_build_tree(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_BUILD); }
int bh_summarize(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_SUMMARIZE); }
int bh_sort(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_SORT); }
int bh_calculate_force(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_FORCE); }
int bh_integrate(bh_sim *sim) { BH_ENTER(sim); return singleStage(s, BH_STAGE_INTEGRATE); }

int bh_stage_async(bh_sim *sim, int32_t stage) { BH_ENTER(sim); return launchStage(s, stage); }

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
    if (first < 0 || count < 0 || (int64_t)first + count > s->n || (first % s->vote) != 0)
        return fail(s, BH_ERR_ARG, "bad slice [%d, %d): first must be a multiple of vote_width", first, first + count);
    if (count == 0) return BH_OK;
    launchForce(s, first, count, true, false);
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[BH_STAGE_FORCE]++;
    return BH_OK;
}

int bh_apply_acceleration(bh_sim *sim) {
    BH_ENTER(sim);
    bh::apply_acc_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(accPhase(s), s->sorted, s->velacc, s->sc, s->n, s->dt);
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[BH_STAGE_FORCE]++;
    if (s->p2p) s->accPhase ^= 1;  // peers may already store the next step's slices while this kernel still reads
    return BH_OK;
}

void *bh_acc_sorted_device_ptr(bh_sim *sim) { return sim ? S(sim)->accSorted : nullptr; }

int bh_ipc_export(bh_sim *sim, void *handle64) {
    BH_ENTER(sim);
    if (!handle64) return fail(s, BH_ERR_ARG, "handle64 is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    BH_CUDA(s, cudaIpcGetMemHandle(&h, s->accSorted));
    memcpy(handle64, &h, sizeof h);
    return BH_OK;
}

int bh_ipc_set_peers(bh_sim *sim, int32_t nranks, int32_t my_rank, const void *handles) {
    BH_ENTER(sim);
    if (nranks < 1 || nranks > bh::kMaxPeers || my_rank < 0 || my_rank >= nranks || !handles)
        return fail(s, BH_ERR_ARG, "bad peer set (at most %d ranks)", bh::kMaxPeers);
    for (int r = 0; r < nranks; ++r) {
        if (r == my_rank) { s->peerAcc[r] = s->accSorted; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + 64 * (size_t)r, sizeof h);
        void *p = nullptr;
        BH_CUDA(s, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peerAcc[r] = static_cast<float4 *>(p);
    }
    s->nranks = nranks;
    s->rank = my_rank;
    s->p2p = nranks > 1;
    s->accPhase = 0;
    return BH_OK;
}

int bh_calculate_force_slice_p2p(bh_sim *sim, int32_t first, int32_t count) {
    BH_ENTER(sim);
    if (!s->p2p) return fail(s, BH_ERR_ARG, "bh_ipc_set_peers has not been called");
    if (first < 0 || count < 0 || (int64_t)first + count > s->n || (first % s->vote) != 0)
        return fail(s, BH_ERR_ARG, "bad slice [%d, %d): first must be a multiple of vote_width", first, first + count);
    if (count == 0) return BH_OK;
    launchForce(s, first, count, true, false, true);
    BH_CUDA(s, cudaGetLastError());
    s->stageLaunches[BH_STAGE_FORCE]++;
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
    BH_ENTER(sim);
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
    switch (which) {
    case BH_POS_X: case BH_POS_Y: case BH_POS_Z:
        bh::export_node_kernel<<<grid, 256, 0, s->stream>>>(s->node4, which - BH_POS_X, static_cast<float *>(s->staging), count);
        break;
    case BH_MASS:
        bh::export_node_kernel<<<grid, 256, 0, s->stream>>>(s->node4, 3, static_cast<float *>(s->staging), count);
        break;
    case BH_VEL_X: case BH_VEL_Y: case BH_VEL_Z:
        bh::export_velacc_kernel<<<grid, 256, 0, s->stream>>>(s->velacc, 0, which - BH_VEL_X, static_cast<float *>(s->staging), s->n, count);
        break;
    case BH_ACC_X: case BH_ACC_Y: case BH_ACC_Z:
        bh::export_velacc_kernel<<<grid, 256, 0, s->stream>>>(s->velacc, 1, which - BH_ACC_X, static_cast<float *>(s->staging), s->n, count);
        break;
    case BH_BODY_COUNT:
        bh::export_shifted_kernel<<<grid, 256, 0, s->stream>>>(s->count, s->n, static_cast<int *>(s->staging), count);
        break;
    case BH_START:
        bh::export_shifted_kernel<<<grid, 256, 0, s->stream>>>(s->start, s->n, static_cast<int *>(s->staging), count);
        break;
    case BH_CHILD:
        bh::export_shifted_kernel<<<grid, 256, 0, s->stream>>>(s->child, 8 * (long long)s->n, static_cast<int *>(s->staging), count);
        break;
    case BH_SORTED: {
        const int64_t head = std::min<int64_t>(count, s->n);
        BH_CUDA(s, cudaMemcpyAsync(s->staging, s->sorted, 4 * (size_t)head, cudaMemcpyDeviceToDevice, s->stream));
        if (count > head) BH_CUDA(s, cudaMemsetAsync(static_cast<int *>(s->staging) + head, 0, 4 * (size_t)(count - head), s->stream));
        break;
    }
    default:
        return fail(s, BH_ERR_ARG, "unknown buffer %d", which);
    }
    BH_CUDA(s, cudaGetLastError());
    BH_CUDA(s, cudaMemcpyAsync(dst, s->staging, 4 * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    return BH_OK;
}

int bh_copy_vertices(bh_sim *sim, float *pos4, float *vel4) {
    BH_ENTER(sim);
    if (!pos4 && !vel4) return BH_OK;
    const size_t bytes = sizeof(float4) * (size_t)s->n;
    int rc = ensureStaging(s, 2 * bytes);
    if (rc) return rc;
    float4 *dp = static_cast<float4 *>(s->staging), *dv = dp + s->n;
    bh::copy_vertices_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->node4, s->velacc, pos4 ? dp : nullptr, vel4 ? dv : nullptr, s->n);
    BH_CUDA(s, cudaGetLastError());
    if (pos4) BH_CUDA(s, cudaMemcpyAsync(pos4, dp, bytes, cudaMemcpyDeviceToHost, s->stream));
    if (vel4) BH_CUDA(s, cudaMemcpyAsync(vel4, dv, bytes, cudaMemcpyDeviceToHost, s->stream));
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    return BH_OK;
}

int bh_stats(bh_sim *sim, bh_stats_t *out) {
    BH_ENTER(sim);
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
    for (int i = 0; i < BH_NUM_STAGES; ++i) {
        out->stage_ms[i] = s->stageMs[i];
        out->stage_launches[i] = s->stageLaunches[i];
    }
    out->interactions = (int64_t)s->hostSc->interactions;
    out->opens = (int64_t)s->hostSc->opens;
    return BH_OK;
}

int bh_reset_stats(bh_sim *sim) {
    BH_ENTER(sim);
    for (int i = 0; i < BH_NUM_STAGES; ++i) {
        s->stageMs[i] = 0.0;
        s->stageLaunches[i] = 0;
    }
    s->stepsTimed = 0;
    return BH_OK;
}

int32_t bh_number_of_bodies(bh_sim *sim) { return sim ? S(sim)->n : BH_ERR_ARG; }

int bh_generate_universe(bh_sim *sim, int32_t kind, uint64_t seed, float p0, float p1, float p2) {
    BH_ENTER(sim);
    if (kind < 0 || kind > 2) return fail(s, BH_ERR_ARG, "unknown universe kind %d", kind);
    int rc = resetState(s);
    if (rc) return rc;
    bh::generate_kernel<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->node4, s->velacc, s->sorted, s->n, kind, seed, p0, p1, p2);
    BH_CUDA(s, cudaGetLastError());
    BH_CUDA(s, cudaStreamSynchronize(s->stream));
    return BH_OK;
}

int bh_diagnostics(bh_sim *sim, int32_t with_potential, bh_diag_t *out) {
    BH_ENTER(sim);
    if (!out) return fail(s, BH_ERR_ARG, "out is NULL");
    int rc = ensureStaging(s, 8 * sizeof(double));
    if (rc) return rc;
    double *d = static_cast<double *>(s->staging);
    BH_CUDA(s, cudaMemsetAsync(d, 0, 8 * sizeof(double), s->stream));
    bh::kinetic_kernel<<<std::min((s->n + 255) / 256, s->numSMs * 8), 256, 0, s->stream>>>(s->node4, s->velacc, d, s->n);
    if (with_potential)
        bh::potential_kernel<<<(s->n + bh::kPotTile - 1) / bh::kPotTile, bh::kPotTile, 0, s->stream>>>(s->node4, d, s->n, s->eps);
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
struct UniverseFile {
    int32_t n = 0;
    std::string error;
    float *arrays[7] = {};
    ~UniverseFile() { for (auto *a : arrays) free(a); }
};
uint32_t be32(const unsigned char *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
bool readUniverse(const char *path, bool headerOnly, UniverseFile &u) {
    FILE *f = fopen(path, "rb");
    if (!f) { u.error = std::string("cannot open ") + path; return false; }
    unsigned char hdr[10];
    static const unsigned char magic[6] = {0xAC, 0xED, 0x00, 0x05, 0x77, 0x04};  // stream header + TC_BLOCKDATA(4) = writeInt
    if (fread(hdr, 1, 10, f) != 10 || memcmp(hdr, magic, 6) != 0) { fclose(f); u.error = "not a .universe file (Java stream with a leading writeInt expected)"; return false; }
    u.n = (int32_t)be32(hdr + 6);
    if (headerOnly) { fclose(f); return true; }
    static const unsigned char classDesc[] = {0x72, 0x00, 0x02, '[', 'F', 0x0B, 0x9C, 0x81, 0x89, 0x22, 0xE0, 0x0C, 0x42, 0x02, 0x00, 0x00, 0x78, 0x70};
    static const unsigned char classRef[] = {0x71, 0x00, 0x7E, 0x00, 0x00};
    for (int a = 0; a < 7; ++a) {
        unsigned char tag[32];
        if (fread(tag, 1, 2, f) != 2 || tag[0] != 0x75) { fclose(f); u.error = "expected TC_ARRAY"; return false; }
        const bool full = tag[1] == 0x72;
        const size_t rest = (full ? sizeof classDesc : sizeof classRef) - 1;
        if (fread(tag + 2, 1, rest, f) != rest || memcmp(tag + 1, full ? classDesc : classRef, rest + 1) != 0) {
            fclose(f); u.error = "unexpected array class (float[] expected)"; return false;
        }
        unsigned char len[4];
        if (fread(len, 1, 4, f) != 4 || (int32_t)be32(len) != u.n) { fclose(f); u.error = "array length differs from nbodies"; return false; }
        u.arrays[a] = static_cast<float *>(malloc(sizeof(float) * (size_t)std::max(u.n, 1)));
        std::string buf(4 * (size_t)u.n, '\0');
        if (!u.arrays[a] || fread(&buf[0], 1, buf.size(), f) != buf.size()) { fclose(f); u.error = "short read"; return false; }
        for (int32_t i = 0; i < u.n; ++i) {
            const uint32_t v = be32(reinterpret_cast<const unsigned char *>(buf.data()) + 4 * (size_t)i);
            memcpy(&u.arrays[a][i], &v, 4);
        }
    }
    fclose(f);
    return true;
}
}  // namespace

int bh_universe_file_bodies(const char *path, int32_t *nbodies) {
    if (!path || !nbodies) return BH_ERR_ARG;
    UniverseFile u;
    if (!readUniverse(path, true, u)) return fail(nullptr, BH_ERR_ARG, "%s", u.error.c_str());
    *nbodies = u.n;
    return BH_OK;
}

int bh_upload_universe_file(bh_sim *sim, const char *path) {
    BH_ENTER(sim);
    if (!path) return fail(s, BH_ERR_ARG, "path is NULL");
    UniverseFile u;
    if (!readUniverse(path, false, u)) return fail(s, BH_ERR_ARG, "%s", u.error.c_str());
    if (u.n != s->n)  // SerializedUniverseGenerator.java:41-42 IllegalStateException
        return fail(s, BH_ERR_ARG, "invalid amount of bodies for serialized universe (%d in the file, %d in the simulation)", u.n, s->n);
    const float *src[7] = {u.arrays[0], u.arrays[1], u.arrays[2], u.arrays[3], u.arrays[4], u.arrays[5], u.arrays[6]};
    return uploadImpl(s, src, cudaMemcpyHostToDevice);
}

int bh_measure_fp32_peak(int32_t device, double *tflops) {
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
        bh::fp32_peak_kernel<<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *tflops = best;
    return best > 0.0 ? BH_OK : BH_ERR_CUDA;
}

}  // extern "C"
