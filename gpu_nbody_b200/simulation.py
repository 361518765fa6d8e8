"""Host-side mirror of the reference's simulation classes over the C ABI.

The reference host is Java (no JDK in this image), so the host layer above
libbhstep.so is restated in Python with the reference's names and argument
meaning:

* ``AbstractNBodySimulation`` -- simulation/AbstractNBodySimulation.java:5-44
  (``init``, ``initGLBuffers``, ``step``, ``getNumberOfBodies``, ``Mode``).
* ``GPUBarnesHutNBodySimulation`` -- simulation/gpu/GPUBarnesHutNBodySimulation.java
  (GPUBH): ctor ``(mode, nbodies, universeGenerator)`` (:104-108), ``init`` sizes
  the node pool and uploads the generator's output (:111-151), ``step`` runs the
  six stages (:249-271).  The OpenCL context / program / buffer objects of the
  reference are replaced by one ``bh_sim`` handle; every kernel enqueue is one
  C-ABI call.

The Java source of the same binding is in java/ (Panama FFM), see INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib
from .universe import UniverseGenerator


class BhError(RuntimeError):
    """Stands in for the CLException the reference gets from JOCL (GPUBH:40-42)
    and for the IllegalStateException of GPUBH:120 when no device exists."""

    def __init__(self, code, message):
        super().__init__("bhstep error %d: %s" % (code, message))
        self.code = code


class Mode(enum.Enum):
    """AbstractNBodySimulation.java:40-42"""
    GL_INTEROP = 0
    DEFAULT = 1


class AbstractNBodySimulation:
    """simulation/AbstractNBodySimulation.java:5-44"""

    def __init__(self, mode: Mode):
        self.mode = mode

    def step(self):
        raise NotImplementedError

    def initGLBuffers(self, gl, positionVBO, velocityVBO):
        raise NotImplementedError

    def getNumberOfBodies(self):
        raise NotImplementedError

    def init(self, gl):
        raise NotImplementedError


class GPUBarnesHutNBodySimulation(AbstractNBodySimulation):
    """GPUBH, with the reference's compile-time constants as keyword parameters:
    theta (THETA macro = theta**2), eps2 (EPSILON), dt (TIMESTEP), vote_width
    (WARPSIZE = WORKGROUP_SIZE = 16).  Defaults are BASELINE.json's theta = 0.5
    and the shipped EPSILON / TIMESTEP (calculateforce.cl:14,17)."""

    WARPSIZE = 16  # GPUBH:47

    def __init__(self, mode: Mode, nbodies: int, generator: UniverseGenerator, *, theta=0.5, eps2=0.0025,
                 dt=0.025, vote_width=16, device=0, theta_macro=None):
        super().__init__(mode)
        self.nbodies = int(nbodies)
        self.universeGenerator = generator
        self.theta, self.eps2, self.dt = float(theta), float(eps2), float(dt)
        self.vote_width, self.device, self.theta_macro = int(vote_width), int(device), theta_macro
        self._lib = None
        self._sim = C.c_void_p()
        self.numberOfNodes = None
        self._pos4 = self._vel4 = None

    # ---- lifecycle ----------------------------------------------------------------
    def init(self, gl=None):
        """GPUBH:111-151.  ``gl`` is accepted for signature parity; there is no GL
        context sharing (the vertices come back through ``copyVertices``)."""
        self._lib = _lib.load()
        rc = self._lib.bh_create(C.byref(self._sim), self.nbodies, self.theta, self.eps2, self.dt, self.vote_width,
                                 self.device)
        if rc != 0:
            raise BhError(rc, (self._lib.bh_last_error(None) or b"").decode())
        if self.theta_macro is not None:
            self._check(self._lib.bh_set_theta_macro(self._sim, float(self.theta_macro)))
        self.numberOfNodes = int(self._lib.bh_number_of_nodes(self.nbodies))  # GPUBH:130
        if self.universeGenerator is None:  # the caller fills the buffers itself (generateOnDevice, uploadUniverseFile)
            return
        m1 = self.numberOfNodes + 1
        host = [np.zeros(m1, dtype=np.float32) for _ in range(7)]             # GPUBH:134-142
        self.universeGenerator.generate(0, self.nbodies, *host)               # GPUBH:144
        self.upload(*host)                                                    # GPUBH:146 loadBuffers

    def upload(self, x, y, z, vx, vy, vz, mass):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z, vx, vy, vz, mass)]
        if any(a.size < self.nbodies for a in arrs):
            raise ValueError("universe arrays shorter than nbodies")
        self._check(self._lib.bh_upload(self._sim, *(a.ctypes.data for a in arrs)))

    def initGLBuffers(self, gl=None, positionVBO=-1, velocityVBO=-1):
        """GPUBH:230-246 with gl == null: two nbodies*4 float buffers for copyVertices."""
        self._pos4 = np.zeros((self.nbodies, 4), dtype=np.float32)
        self._vel4 = np.zeros((self.nbodies, 4), dtype=np.float32)

    def close(self):
        if self._sim:
            self._lib.bh_destroy(self._sim)
            self._sim = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getNumberOfBodies(self):
        return self.nbodies

    # ---- the step -----------------------------------------------------------------
    def step(self, nsteps: int = 1):
        """GPUBH:249-271 (and :265-266 copyVertices in GL_INTEROP mode)."""
        self._check(self._lib.bh_step(self._sim, nsteps))
        if self.mode == Mode.GL_INTEROP:
            self.copyVertices()

    # executeSimulationKernel(kernel), GPUBH:258-263
    def boundingBox(self): self._check(self._lib.bh_bounding_box(self._sim))
    def buildTree(self): self._check(self._lib.bh_build_tree(self._sim))
    def summarizeTree(self): self._check(self._lib.bh_summarize(self._sim))
    def sort(self): self._check(self._lib.bh_sort(self._sim))
    def calculateForce(self): self._check(self._lib.bh_calculate_force(self._sim))
    def integrate(self): self._check(self._lib.bh_integrate(self._sim))

    def copyVertices(self):
        """copyvertices.cl:8-17; returns (pos4, vel4) host arrays of shape (n, 4)."""
        if self._pos4 is None:
            self.initGLBuffers(None, -1, -1)
        self._check(self._lib.bh_copy_vertices(self._sim, self._pos4.ctypes.data, self._vel4.ctypes.data))
        return self._pos4, self._vel4

    # ---- readBuffer + getData (GPUBH:277-278,294-295,306-312) -----------------------
    def readBuffer(self, name: str, count=None):
        which = _lib.BUFFERS.index(name)
        length = int(self._lib.bh_buffer_length(self._sim, which))
        count = length if count is None else int(count)
        out = np.empty(count, dtype=np.float32 if name in _lib.FLOAT_BUFFERS else np.int32)
        self._check(self._lib.bh_read(self._sim, which, out.ctypes.data, count))
        return out

    def scalar(self, name: str):
        return self.readBuffer(name)[0]

    # ---- options / diagnostics --------------------------------------------------------
    def setProfiling(self, on=True): self._check(self._lib.bh_set_profiling(self._sim, int(on)))
    def setCounting(self, on=True): self._check(self._lib.bh_set_counting(self._sim, int(on)))
    def setGraph(self, on=True): self._check(self._lib.bh_set_graph(self._sim, int(on)))
    def setInsertionOrder(self, mode): self._check(self._lib.bh_set_insertion_order(self._sim, int(mode)))
    def setStream(self, cuda_stream): self._check(self._lib.bh_set_stream(self._sim, C.c_void_p(cuda_stream)))
    def resetStats(self): self._check(self._lib.bh_reset_stats(self._sim))

    def diagnostics(self, with_potential=True):
        """printEnergy / printImpulse (GPUBH:305-365) on the device: dict(ekin, epot, etot, px, py, pz, mass).
        with_potential: True / 1 = exact tiled O(N^2) sum, 2 = through the tree (one more force walk), False = none."""
        d = _lib.BhDiag()
        self._check(self._lib.bh_diagnostics(self._sim, int(with_potential), C.byref(d)))
        out = {k: getattr(d, k) for k, _ in d._fields_}
        out["etot"] = out["ekin"] + out["epot"]
        return out

    def generateOnDevice(self, kind, seed, p0=0.0, p1=0.0, p2=0.0):
        """Seeded generator on the device: kind "cubic" (p0 = range), "plummer", "disk" (r, velocityMultiplier, centerMass)."""
        k = {"cubic": 0, "plummer": 1, "disk": 2}[kind]
        self._check(self._lib.bh_generate_universe(self._sim, k, int(seed), float(p0), float(p1), float(p2)))

    def setForceDeepWalk(self, on=True):
        """Validation / A-B timing: run the force stage with the shared-stack walk kernel (the one used for 32-wide votes)."""
        self._check(self._lib.bh_set_force_deep_walk(self._sim, int(on)))

    def setVertexBuffers(self, pos4_device_ptr, vel4_device_ptr):
        """GL_INTEROP (GPUBH:230-246,265-266): device pointers of two float4[nbodies] buffers (e.g. CUDA-mapped GL
        vertex buffers) that every step's finish pass fills; (0, 0) switches it off."""
        self._check(self._lib.bh_set_vertex_buffers(self._sim, C.c_void_p(pos4_device_ptr or None), C.c_void_p(vel4_device_ptr or None)))

    def copyVerticesDevice(self, pos4_device_ptr, vel4_device_ptr):
        """copyvertices.cl:8-17 into device buffers, asynchronously on the simulation's stream."""
        self._check(self._lib.bh_copy_vertices_device(self._sim, C.c_void_p(pos4_device_ptr or None), C.c_void_p(vel4_device_ptr or None)))

    def writeUniverseFile(self, path):
        """UniverseSerializer.serialize (UniverseSerializer.java:25-34) of the current state, natively."""
        self._check(self._lib.bh_write_universe_file(self._sim, str(path).encode()))

    def uploadUniverseFile(self, path):
        """SerializedUniverseGenerator without a JVM: read a .universe file natively and upload it."""
        self._check(self._lib.bh_upload_universe_file(self._sim, str(path).encode()))

    def stats(self):
        st = _lib.BhStats()
        self._check(self._lib.bh_stats(self._sim, C.byref(st)))
        d = {k: getattr(st, k) for k, _ in st._fields_ if k not in ("stage_ms", "stage_launches")}
        d["stage_ms"] = dict(zip(_lib.STAGES, list(st.stage_ms)))
        d["stage_launches"] = dict(zip(_lib.STAGES, list(st.stage_launches)))
        return d

    @property
    def handle(self):
        return self._sim

    def _check(self, rc):
        if rc != 0:
            raise BhError(rc, (self._lib.bh_last_error(self._sim) or b"").decode())
