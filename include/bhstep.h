/*
 * bhstep.h -- C ABI of libbhstep.so, the B200 (sm_100a) Barnes-Hut step.
 *
 * This is the drop-in boundary for the hot path of bneukom/gpu-nbody: every
 * entry point replaces a JOCL/OOCL call sequence of
 *   src/ch/fhnw/woipv/nbody/simulation/gpu/GPUBarnesHutNBodySimulation.java
 * (cited as GPUBH below; kernel sources under kernels/nbody/).  Plain C types
 * only; no torch, no CUDA types (a stream travels as void*).  All functions
 * return 0 on success, a negative bh_status on CUDA/argument failure (text via
 * bh_last_error) and a positive value when the device `error` buffer is set
 * (1 = cell pool exhausted or tree deeper than 64 levels, exactly the two
 * conditions of buildtree.cl:112-119 and calculateforce.cl:69-73;
 * 2 = a device-side wait exceeded its spin budget; 3 = the multi-GPU peer
 * barrier timed out).
 *
 * Residency: the sort stage (kernels/nbody/sort.cl) waits, thread on thread, for
 * the `start` value of a cell's parent, as the reference does.  It is launched
 * cooperatively (cudaLaunchCooperativeKernel): all its CTAs are co-resident or
 * the launch waits for the device, so work on other streams of the same
 * device (another bh_sim, NCCL, MPS clients) can delay it but not starve it.
 * Every device-side wait is bounded (2^24 polls, then error 2), never a hang.
 *
 * Threading: a bh_sim is thread-compatible, not thread-safe (the reference is
 * driven from one JOGL animator thread, NBodyVisualizer.java:460-470).  Every
 * call selects the simulation's device first, so JVM callers may hop threads.
 *
 * There is no CPU fallback: bh_create fails with BH_ERR_NO_DEVICE when no
 * CUDA device is present.
 */
#ifndef BHSTEP_H
#define BHSTEP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bh_sim bh_sim; /* opaque; one per simulation and device */

enum bh_status {
    BH_OK = 0,
    BH_ERR_CUDA = -1,      /* a CUDA runtime call failed                    */
    BH_ERR_ARG = -2,       /* bad argument                                  */
    BH_ERR_NO_DEVICE = -3, /* no CUDA device (GPUBH:120 IllegalStateException) */
    BH_ERR_ALLOC = -4
};

/* Logical buffers in the reference's kernel-argument order (GPUBH:198-205). */
enum bh_buffer {
    BH_POS_X = 0, BH_POS_Y, BH_POS_Z,     /* float[M+1]  bodies 0..N-1, cells N..M (root M) */
    BH_VEL_X, BH_VEL_Y, BH_VEL_Z,         /* float[M+1]  */
    BH_ACC_X, BH_ACC_Y, BH_ACC_Z,         /* float[M+1]  */
    BH_STEP,                              /* int[1]      */
    BH_BLOCK_COUNT,                       /* int[1]      */
    BH_BODY_COUNT,                        /* int[M+1]    */
    BH_RADIUS,                            /* float[1]    */
    BH_MAX_DEPTH,                         /* int[1]      */
    BH_BOTTOM,                            /* int[1]      */
    BH_MASS,                              /* float[M+1]  */
    BH_CHILD,                             /* int[8(M+1)] */
    BH_START,                             /* int[M+1]    */
    BH_SORTED,                            /* int[M+1]    */
    BH_ERROR,                             /* int[1]      */
    BH_NUM_BUFFERS
};

enum bh_stage { BH_STAGE_BBOX = 0, BH_STAGE_BUILD, BH_STAGE_SUMMARIZE, BH_STAGE_SORT, BH_STAGE_FORCE,
                BH_STAGE_INTEGRATE, BH_NUM_STAGES };

typedef struct bh_stats_t {
    int32_t nbodies;          /* N */
    int32_t number_of_nodes;  /* M  (GPUBH:219-227) */
    int32_t cells_used;       /* M - bottom + 1 after the last build */
    int32_t max_depth;        /* running maximum, as in the reference */
    int32_t step;             /* value of the `step` buffer */
    int32_t error;            /* value of the `error` buffer */
    int64_t steps_timed;      /* stage executions accumulated in stage_ms (profiling on) */
    double stage_ms[BH_NUM_STAGES];      /* summed CUDA-event time per stage */
    int64_t stage_launches[BH_NUM_STAGES]; /* kernel launches per stage since reset */
    int64_t interactions;     /* (body,node) force evaluations of the last counted force call */
    int64_t opens;            /* (body,cell) opening tests that pushed, same call */
    double barrier_ms;        /* summed CUDA-event time of the multi-GPU peer barrier (ABI >= 2) */
    int32_t deep_walk;        /* 1 = the force stage runs the shared-stack walk kernel (vote width 32 or bh_set_force_deep_walk) */
    int32_t walk_spills;      /* times a vote group's cell stack spilled to global memory in the last force walk (deep trees) */
} bh_stats_t;

/* GPUBH.init():111-151 -- context, node-pool sizing (219-227), buffer creation (153-181).
 * theta is the opening angle (the reference's THETA macro is theta^2,
 * calculateforce.cl:15-16); eps2 is EPSILON (added to r^2); dt is TIMESTEP;
 * vote_width is the reference's WARPSIZE/WORKGROUP_SIZE (16 = parity; 32 = one
 * vote per hardware warp, not reference-exact). */
int bh_create(bh_sim **out, int32_t nbodies, float theta, float eps2, float dt, int32_t vote_width, int32_t device);
void bh_destroy(bh_sim *sim);
const char *bh_last_error(bh_sim *sim); /* sim may be NULL: error of the last failed bh_create */

/* Override theta^2 directly (e.g. the shipped THETA (1.5f), calculateforce.cl:16). */
int bh_set_theta_macro(bh_sim *sim, float theta_macro);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL = CUDA's default
 * stream).  A new simulation runs on a private non-blocking stream; bh_use_private_stream
 * returns to it. */
int bh_set_stream(bh_sim *sim, void *cuda_stream);
int bh_use_private_stream(bh_sim *sim);
/* 1 = record CUDA events around every stage (bh_stats.stage_ms; steps are then launched kernel by kernel instead of
 * replaying the step's CUDA graph); 0 = off (default).  Events exist for 64 steps between two calls that wait for the
 * device: bh_step splits longer runs itself, bh_step_async times the first 64 steps of a longer burst only. */
int bh_set_profiling(bh_sim *sim, int32_t on);
/* 1 = count interactions/opens in the next force calls (slower kernel variant); 0 = off. */
int bh_set_counting(bh_sim *sim, int32_t on);
/* Body storage / insertion order: 1 (default) = after every step the bodies are physically stored in
 * that step's sorted (DFS / Morton-like) order, which is the next build's insertion order (coalesced
 * loads); 0 = bodies stay in upload order.  Results and everything bh_read returns are identical. */
int bh_set_insertion_order(bh_sim *sim, int32_t mode);
/* Validation / A-B timing: 1 = run the force stage with the shared-stack walk kernel (the one used for
 * 32-wide votes) instead of the per-group walk; 0 = default.  Same results up to summation order. */
int bh_set_force_deep_walk(bh_sim *sim, int32_t on);
/* 1 (default) = bh_step replays one captured CUDA graph per step (six kernels, fixed arguments); 0 = six launches. */
int bh_set_graph(bh_sim *sim, int32_t on);

/* createBuffer(CL_MEM_COPY_HOST_PTR, ...) for the seven generator outputs (GPUBH:155-170):
 * caller-owned host SoA arrays of length nbodies are copied; all other buffers are
 * reset to their initial values (step=-1, maxDepth=1, rest 0).  Where the bodies are stored is internal: an
 * upload over an existing state keeps body i in the slot the previous body i had (the last step's tree order), so
 * a host that sends its bodies every step does not lose the tree stages' locality.  Nothing bh_read or
 * bh_copy_vertices returns depends on it. */
int bh_upload(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
              const float *vz, const float *mass);
/* The same without waiting, for pinned host arrays: positions and masses are copied on the simulation's stream, the
 * velocities on a second one, and the next bh_step starts its tree stages and force walk as soon as the former have
 * arrived (only its last pass needs the velocities).  The arrays must stay valid and unchanged until a call that
 * waits for the device returns (bh_step, bh_check, bh_read ...). */
int bh_upload_async(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                    const float *vz, const float *mass);
/* Same with device pointers (inputs already resident in HBM). */
int bh_upload_device(bh_sim *sim, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                     const float *vz, const float *mass);

/* The per-stage kernel contract, one call per executeSimulationKernel (GPUBH:258-263,273-275).
 * Each enqueues on the simulation's stream and waits for it, like finish() at GPUBH:275.
 * A stage whose inputs do not exist yet is refused with BH_ERR_ARG instead of reading uninitialised device memory:
 * build_tree needs bounding_box, summarize needs build_tree, sort needs build_tree + summarize, calculate_force needs
 * summarize + sort (each since the last upload); a calculate_force over the previous step's tree is allowed, as in the
 * reference, but summarize / sort are not once a step has physically reordered the bodies (the tree names them by slot). */
int bh_bounding_box(bh_sim *sim);    /* kernels/nbody/boundingbox.cl:21     */
int bh_build_tree(bh_sim *sim);      /* kernels/nbody/buildtree.cl:13       */
int bh_summarize(bh_sim *sim);       /* kernels/nbody/summarizetree.cl:16   */
int bh_sort(bh_sim *sim);            /* kernels/nbody/sort.cl:13            */
int bh_calculate_force(bh_sim *sim); /* kernels/nbody/calculateforce.cl:26  */
int bh_integrate(bh_sim *sim);       /* kernels/nbody/integrate.cl:18       */

/* GPUBH.step():249-271, nsteps times: the six stages in order, no host sync
 * between them, one sync and one look at the error buffer at the end. */
int bh_step(bh_sim *sim, int32_t nsteps);
/* Same without the final sync (caller synchronises its stream, then bh_check). */
int bh_step_async(bh_sim *sim, int32_t nsteps);
int bh_check(bh_sim *sim); /* sync + error buffer */

/* Multi-GPU slice contract (no counterpart in the single-device reference).
 * bh_calculate_force_slice walks the tree for sorted slots [first, first+count)
 * (first a multiple of vote_width) and stores float4 {ax,ay,az,0} per slot into
 * the sorted-order acceleration buffer; after the caller has all-gathered that
 * buffer across ranks, bh_finish_async performs the velocity correction and acc
 * store of calculateforce.cl:174-185 and the integrate of integrate.cl for all N
 * bodies in one pass (bh_apply_acceleration: the velocity correction alone).  All async. */
int bh_calculate_force_slice(bh_sim *sim, int32_t first, int32_t count);
int bh_apply_acceleration(bh_sim *sim);
int bh_finish_async(bh_sim *sim);
void *bh_acc_sorted_device_ptr(bh_sim *sim); /* float4[N + 2048] in device memory (slack for equal, aligned slices) */
/* Peer-memory variant (one process per GPU, one node): the all-gather is fused into the force kernel and the
 * whole sliced step runs inside bh_step / bh_step_async (one CUDA graph per step, no host or NCCL call in it).
 * Every rank exports its acceleration buffer (bh_ipc_export: 64-byte CUDA IPC handle), the host exchanges
 * the handles (any transport) and hands all of them to bh_ipc_set_peers; bh_set_slice names the rank's
 * slice of the sorted order.  From then on the step's force stage walks the slice and stores each slot's
 * result into the own buffer and, over NVLink, into every peer's; a device-side barrier (flags in the same
 * peer-mapped allocation, release/acquire at system scope) separates it from the finish pass, which every
 * rank runs over all N bodies.  The buffer is double-buffered per step, so one barrier per step is enough.
 * All ranks must call bh_step the same number of times.  bh_ipc_clear_peers unmaps the peers and returns to
 * the single-GPU step (call it on every rank if bh_ipc_set_peers failed on any).
 * bh_calculate_force_slice_p2p / bh_peer_barrier are the same two pieces as single async calls. */
int bh_ipc_export(bh_sim *sim, void *handle64);
int bh_ipc_set_peers(bh_sim *sim, int32_t nranks, int32_t my_rank, const void *handles /* nranks x 64 bytes */);
int bh_ipc_clear_peers(bh_sim *sim);
int bh_set_slice(bh_sim *sim, int32_t first, int32_t count);
int bh_calculate_force_slice_p2p(bh_sim *sim, int32_t first, int32_t count);
int bh_peer_barrier(bh_sim *sim);
/* Async single stages for callers that sequence their own stream. */
int bh_stage_async(bh_sim *sim, int32_t stage /* enum bh_stage */);

/* queue.readBuffer(mem) + mem.getData() (GPUBH:277-278,294-295,306-312): copies the first
 * `count` elements of logical buffer `which` to host memory, in the reference's
 * index conventions whatever the internal layout. */
int bh_read(bh_sim *sim, int32_t which, void *dst, int64_t count);
int64_t bh_buffer_length(bh_sim *sim, int32_t which);

/* kernels/nbody/copyvertices.cl:8-17 with host destinations: pos4[i] = {x,y,z,1},
 * vel4[i] = {vx,vy,vz,1}, i < nbodies.  Either pointer may be NULL. */
int bh_copy_vertices(bh_sim *sim, float *pos4, float *vel4);
/* The same without waiting: the vertices are exported on the simulation's stream and copied to the (pinned) host buffers
 * on a second stream, so the read-back overlaps whatever the caller enqueues next -- the next bh_upload (other PCIe
 * direction), the next bh_step.  The host buffers are valid after bh_wait_copies (or the next bh_copy_vertices*). */
int bh_copy_vertices_async(bh_sim *sim, float *pos4, float *vel4);
int bh_wait_copies(bh_sim *sim);
/* The same with DEVICE destinations (float4[nbodies] each) -- what a CUDA-mapped OpenGL vertex buffer is
 * (cudaGraphicsResourceGetMappedPointer; GPUBH:230-246 createFromGLBuffer + :253-256 acquire): async on the
 * simulation's stream, no host staging. */
int bh_copy_vertices_device(bh_sim *sim, void *pos4_device, void *vel4_device);
/* GL_INTEROP mode of GPUBH.step() (:265-266): from now on every step's finish pass also writes the float4
 * vertices into these device buffers (fused, no extra kernel); NULL, NULL switches it off. */
int bh_set_vertex_buffers(bh_sim *sim, void *pos4_device, void *vel4_device);

/* Seeded universe generators on the device (the reference's universe generators draw from the unseeded
 * Math.random() on the host and upload): Philox4x32-10, one subsequence per body; resets the other buffers
 * like bh_upload.  kind 0 = RandomCubicUniverseGenerator(range = p0) (RandomCubicUniverseGenerator.java:13-17),
 * 1 = PlummerUniverseGenerator (PlummerUniverseGenerator.java:8-41), 2 = RotatingDiskGalaxyGenerator(r = p0,
 * velocityMultiplier = p1, centerMass = p2) (RotatingDiskGalaxyGenerator.java:17-43). */
enum bh_universe_kind { BH_UNIVERSE_RANDOM_CUBIC = 0, BH_UNIVERSE_PLUMMER = 1, BH_UNIVERSE_ROTATING_DISK = 2 };
int bh_generate_universe(bh_sim *sim, int32_t kind, uint64_t seed, float p0, float p1, float p2);

/* printEnergy / printImpulse (GPUBH:305-365) as device reductions instead of O(N^2) host loops:
 * kinetic energy, momentum, total mass, and the softened potential -sum_{i<j} m_i m_j / sqrt(r^2 + eps2)
 * (the parity tests' energy formula; the reference's own printEnergy uses an unsoftened, doubled potential
 * and is not reproduced): with_potential = 1 by a tiled direct sum (exact, O(N^2): up to ~10^5 bodies),
 * = 2 through the tree (the force walk's opening rule applied to m / r instead of m d / r^3: one more walk,
 * ~1e-3 relative; the state is not advanced), = 0 not at all. */
typedef struct bh_diag_t { double ekin, epot, px, py, pz, mass; } bh_diag_t;
int bh_diagnostics(bh_sim *sim, int32_t with_potential, bh_diag_t *out);

/* SerializedUniverseGenerator (universe/serialize/SerializedUniverseGenerator.java:21-53) without a JVM:
 * reads a .universe file (Java ObjectOutputStream layout of UniverseSerializer.java:25-34) and uploads it;
 * fails like the reference when the body count differs from the simulation's. */
int bh_universe_file_bodies(const char *path, int32_t *nbodies);
int bh_upload_universe_file(bh_sim *sim, const char *path);
/* The loader alone (no device needed): fills caller-owned host arrays of `capacity` floats each. */
int bh_read_universe_file(const char *path, int32_t capacity, float *x, float *y, float *z, float *vx, float *vy,
                          float *vz, float *mass);
/* UniverseSerializer.serialize (universe/serialize/UniverseSerializer.java:25-34) for the simulation's CURRENT
 * state, host numbering: a state dump every SerializedUniverseGenerator (Java or bh_upload_universe_file) reads. */
int bh_write_universe_file(bh_sim *sim, const char *path);

int bh_stats(bh_sim *sim, bh_stats_t *out);
int bh_reset_stats(bh_sim *sim);
int32_t bh_number_of_bodies(bh_sim *sim);          /* getNumberOfBodies(), GPUBH:377-379 */
int32_t bh_number_of_nodes(int32_t nbodies);       /* calculateNumberOfNodes, GPUBH:219-227 */
int32_t bh_abi_version(void);
/* Measurement utility (no reference counterpart): sustained FP32 FMA rate of `device`
 * in TFLOP/s, the roofline denominator bench.py uses for the force kernel. */
int bh_measure_fp32_peak(int32_t device, double *tflops);
/* The same for packed FFMA2 instructions with three distinct register-pair operands (the walk's instruction): what the
 * register file lets the fp32 pipe sustain without constant / immediate operands.  Reported beside the roofline, not used as its peak. */
int bh_measure_fp32x2_rate(int32_t device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* BHSTEP_H */
