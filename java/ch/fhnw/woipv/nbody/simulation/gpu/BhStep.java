package ch.fhnw.woipv.nbody.simulation.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (JDK 22+) binding of include/bhstep.h -- what replaces the
 * net.benjaminneukom.oocl.cl.* / org.jocl.* calls of GPUBarnesHutNBodySimulation
 * (CLPlatform/CLContext/CLCommandQueue/CLKernel/CLMemory, GPUBH:120-122,155-190,199,266-277).
 *
 * NOT COMPILED in this repository's build image (no JDK there); it mirrors, call
 * for call, gpu_nbody_b200/_lib.py, which is exercised by the test-suite.
 */
public final class BhStep implements AutoCloseable {
	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("bhstep.library", "libbhstep.so"), Arena.global());

	private static MethodHandle fn(final String name, final FunctionDescriptor d) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
	}

	private static final MethodHandle CREATE = fn("bh_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_FLOAT, JAVA_FLOAT, JAVA_FLOAT, JAVA_INT, JAVA_INT));
	private static final MethodHandle DESTROY = fn("bh_destroy", FunctionDescriptor.ofVoid(ADDRESS));
	private static final MethodHandle LAST_ERROR = fn("bh_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
	private static final MethodHandle UPLOAD = fn("bh_upload", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	private static final MethodHandle STEP = fn("bh_step", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
	private static final MethodHandle READ = fn("bh_read", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
	private static final MethodHandle COPY_VERTICES = fn("bh_copy_vertices", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
	private static final MethodHandle NUMBER_OF_NODES = fn("bh_number_of_nodes", FunctionDescriptor.of(JAVA_INT, JAVA_INT));
	private static final MethodHandle SET_THETA_MACRO = fn("bh_set_theta_macro", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_FLOAT));
	private static final MethodHandle SET_VERTEX_BUFFERS = fn("bh_set_vertex_buffers", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
	private static final MethodHandle WRITE_UNIVERSE = fn("bh_write_universe_file", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	private static final String[] STAGES = { "bh_bounding_box", "bh_build_tree", "bh_summarize", "bh_sort", "bh_calculate_force", "bh_integrate" };
	private static final MethodHandle[] STAGE = new MethodHandle[STAGES.length];
	static {
		for (int i = 0; i < STAGES.length; ++i)
			STAGE[i] = fn(STAGES[i], FunctionDescriptor.of(JAVA_INT, ADDRESS));
	}

	/** Buffer ids in the kernel-argument order of GPUBH:198-205 (enum bh_buffer). */
	public static final int POS_X = 0, POS_Y = 1, POS_Z = 2, VEL_X = 3, VEL_Y = 4, VEL_Z = 5, ACC_X = 6, ACC_Y = 7, ACC_Z = 8, STEP_BUF = 9,
			BLOCK_COUNT = 10, BODY_COUNT = 11, RADIUS = 12, MAX_DEPTH = 13, BOTTOM = 14, MASS = 15, CHILD = 16, START = 17, SORTED = 18, ERROR = 19;

	private final Arena arena = Arena.ofShared();
	private MemorySegment sim;
	private final int nbodies;

	/** Replaces CLPlatform.getFirst().getDevice(...), createContext, createCommandQueue, loadKernels (GPUBH:120-122,183-192). */
	public BhStep(final int nbodies, final float theta, final float eps2, final float dt, final int voteWidth, final int device) {
		this.nbodies = nbodies;
		try {
			final MemorySegment out = arena.allocate(ADDRESS);
			final int rc = (int) CREATE.invokeExact(out, nbodies, theta, eps2, dt, voteWidth, device);
			if (rc != 0)
				throw new IllegalStateException("bh_create: " + rc + " " + message(MemorySegment.NULL)); // GPUBH:120 orElseThrow
			sim = out.get(ADDRESS, 0);
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	public static int numberOfNodes(final int nbodies) {
		try {
			return (int) NUMBER_OF_NODES.invokeExact(nbodies);
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** Replaces the seven createBuffer(CL_MEM_COPY_HOST_PTR, float[]) calls of loadBuffers (GPUBH:155-170). */
	public void upload(final float[] x, final float[] y, final float[] z, final float[] vx, final float[] vy, final float[] vz, final float[] mass) {
		try (Arena a = Arena.ofConfined()) {
			final float[][] src = { x, y, z, vx, vy, vz, mass };
			for (final float[] arr : src)
				if (arr == null || arr.length < nbodies)
					throw new IllegalArgumentException("universe arrays must hold at least nbodies elements");
			final MemorySegment[] seg = new MemorySegment[7];
			for (int i = 0; i < 7; ++i) {
				seg[i] = a.allocate(JAVA_FLOAT, nbodies);
				MemorySegment.copy(src[i], 0, seg[i], JAVA_FLOAT, 0, nbodies);
			}
			check((int) UPLOAD.invokeExact(sim, seg[0], seg[1], seg[2], seg[3], seg[4], seg[5], seg[6]));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** Replaces the six executeSimulationKernel calls + finish() of step() (GPUBH:258-268). */
	public void step(final int nsteps) {
		try {
			check((int) STEP.invokeExact(sim, nsteps));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** One executeSimulationKernel(kernel) (GPUBH:273-275); stage 0..5 in step() order. */
	public void stage(final int stage) {
		try {
			check((int) STAGE[stage].invokeExact(sim));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** Replaces commandQueue.readBuffer(mem) + mem.getData() (GPUBH:277-278,294-295,306-312). */
	public float[] readFloats(final int which, final int count) {
		if (count < 0)
			throw new IllegalArgumentException("count < 0");
		try (Arena a = Arena.ofConfined()) {
			final MemorySegment dst = a.allocate(JAVA_FLOAT, count);
			check((int) READ.invokeExact(sim, which, dst, (long) count));
			return dst.toArray(JAVA_FLOAT);
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	public int[] readInts(final int which, final int count) {
		if (count < 0)
			throw new IllegalArgumentException("count < 0");
		try (Arena a = Arena.ofConfined()) {
			final MemorySegment dst = a.allocate(JAVA_INT, count);
			check((int) READ.invokeExact(sim, which, dst, (long) count));
			return dst.toArray(JAVA_INT);
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** The THETA macro of calculateforce.cl:15-16 (= theta^2), e.g. the shipped 1.5f. */
	public void setThetaMacro(final float thetaMacro) {
		try {
			check((int) SET_THETA_MACRO.invokeExact(sim, thetaMacro));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** GL_INTEROP without host staging: DEVICE addresses of two float4[nbodies] buffers (CUDA-mapped GL vertex buffers). */
	public void setVertexBuffers(final long pos4Device, final long vel4Device) {
		try {
			check((int) SET_VERTEX_BUFFERS.invokeExact(sim, MemorySegment.ofAddress(pos4Device), MemorySegment.ofAddress(vel4Device)));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** UniverseSerializer.serialize of the current device state (UniverseSerializer.java:25-34). */
	public void writeUniverseFile(final String path) {
		try (Arena a = Arena.ofConfined()) {
			check((int) WRITE_UNIVERSE.invokeExact(sim, a.allocateFrom(path)));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** Replaces the copyVertices kernel launch (GPUBH:265-266) for direct NIO buffers of nbodies*4 floats each (either may be null). */
	public void copyVertices(final java.nio.ByteBuffer pos4, final java.nio.ByteBuffer vel4) {
		if ((pos4 != null && (!pos4.isDirect() || pos4.capacity() < 16L * nbodies)) || (vel4 != null && (!vel4.isDirect() || vel4.capacity() < 16L * nbodies)))
			throw new IllegalArgumentException("copyVertices needs direct buffers of nbodies * 16 bytes");
		copyVertices(pos4 == null ? MemorySegment.NULL : MemorySegment.ofBuffer(pos4), vel4 == null ? MemorySegment.NULL : MemorySegment.ofBuffer(vel4));
	}

	/** Replaces the copyVertices kernel launch (GPUBH:265-266); pos4/vel4 are nbodies*4 floats (mapped GL buffers or heap). */
	public void copyVertices(final MemorySegment pos4, final MemorySegment vel4) {
		try {
			check((int) COPY_VERTICES.invokeExact(sim, pos4, vel4));
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	private void check(final int rc) {
		if (rc != 0)
			throw new IllegalStateException("bhstep error " + rc + ": " + message(sim)); // stands in for JOCL's CLException (GPUBH:40-42)
	}

	private static String message(final MemorySegment s) {
		try {
			final MemorySegment p = (MemorySegment) LAST_ERROR.invokeExact(s);
			return p.equals(MemorySegment.NULL) ? "" : p.reinterpret(512).getString(0);
		} catch (final Throwable t) {
			return "";
		}
	}

	@Override
	public void close() {
		try {
			if (sim != null)
				DESTROY.invokeExact(sim);
		} catch (final Throwable ignored) {
		}
		sim = null;
		arena.close();
	}
}
