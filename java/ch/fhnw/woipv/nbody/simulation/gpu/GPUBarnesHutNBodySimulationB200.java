package ch.fhnw.woipv.nbody.simulation.gpu;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

import com.jogamp.opengl.GL;
import com.jogamp.opengl.GL3;

import ch.fhnw.woipv.nbody.simulation.AbstractNBodySimulation;
import ch.fhnw.woipv.nbody.simulation.universe.UniverseGenerator;

/**
 * Drop-in for GPUBarnesHutNBodySimulation: same constructor and the four
 * AbstractNBodySimulation methods, the OpenCL calls replaced by libbhstep.so.
 * NOT COMPILED in this repository's build image (no JDK, no JOGL jars on the class path there).
 * In NBodyVisualizer.java:208-221 replace `new GPUBarnesHutNBodySimulation(mode, n, generator)` by
 * `new GPUBarnesHutNBodySimulationB200(mode, n, generator)`.
 *
 * GL_INTEROP (GPUBH:230-246, 253-256, 265-266): the reference's copyVertices kernel writes the two vertex buffers
 * through createFromGLBuffer.  Here the vertices are produced on the device and reach the VBOs in one of two ways:
 * (a) portable: bh_copy_vertices into direct host buffers + glBufferSubData (what this class does);
 * (b) zero-copy: register the VBOs with CUDA (cudaGraphicsGLRegisterBuffer / cudaGraphicsMapResources /
 *     cudaGraphicsResourceGetMappedPointer, e.g. through JCuda) and hand the mapped DEVICE pointers to
 *     bh_set_vertex_buffers once: every step's finish pass then writes the float4 vertices itself.
 */
public class GPUBarnesHutNBodySimulationB200 extends AbstractNBodySimulation {
	/** the shipped calculateforce.cl:16 has THETA (1.5f); its commented-out line 15 is 0.5f * 0.5f */
	public static final float SHIPPED_THETA_MACRO = 1.5f;

	private final int nbodies;
	private final UniverseGenerator universeGenerator;
	private final float theta;
	private final Float thetaMacro;
	private BhStep bh;
	private GL3 gl;
	private int positionVBO = -1, velocityVBO = -1;
	private ByteBuffer pos4, vel4;

	/** Same trajectories as the shipped reference: THETA macro 1.5f. */
	public GPUBarnesHutNBodySimulationB200(final Mode mode, final int nbodies, final UniverseGenerator generator) {
		this(mode, nbodies, generator, (float) Math.sqrt(SHIPPED_THETA_MACRO), SHIPPED_THETA_MACRO);
	}

	/** theta = opening angle; thetaMacro (may be null) overrides theta^2 exactly, e.g. 1.5f. */
	public GPUBarnesHutNBodySimulationB200(final Mode mode, final int nbodies, final UniverseGenerator generator, final float theta,
			final Float thetaMacro) {
		super(mode);
		this.nbodies = nbodies;
		this.universeGenerator = generator;
		this.theta = theta;
		this.thetaMacro = thetaMacro;
	}

	@Override
	public void init(final GL3 gl) { // GPUBH:111-151
		final int numberOfNodes = BhStep.numberOfNodes(nbodies);
		final float[] x = new float[numberOfNodes + 1], y = new float[numberOfNodes + 1], z = new float[numberOfNodes + 1];
		final float[] vx = new float[numberOfNodes + 1], vy = new float[numberOfNodes + 1], vz = new float[numberOfNodes + 1];
		final float[] mass = new float[numberOfNodes + 1];
		universeGenerator.generate(0, nbodies, x, y, z, vx, vy, vz, mass);
		bh = new BhStep(nbodies, theta, 0.0025f, 0.025f, 16, 0);
		if (thetaMacro != null)
			bh.setThetaMacro(thetaMacro);
		bh.upload(x, y, z, vx, vy, vz, mass);
	}

	@Override
	public void initGLBuffers(final GL3 gl, final int positionVBO, final int velocityVBO) { // GPUBH:230-246
		this.gl = gl;
		this.positionVBO = positionVBO;
		this.velocityVBO = velocityVBO;
		pos4 = ByteBuffer.allocateDirect(16 * nbodies).order(ByteOrder.nativeOrder());
		vel4 = ByteBuffer.allocateDirect(16 * nbodies).order(ByteOrder.nativeOrder());
	}

	@Override
	public void step() { // GPUBH:249-271
		bh.step(1);
		if (mode == Mode.GL_INTEROP) { // GPUBH:265-266: the vertex buffers follow every step
			if (pos4 == null)
				throw new IllegalStateException("initGLBuffers has not been called");
			bh.copyVertices(pos4, vel4);
			if (gl != null) {
				gl.glBindBuffer(GL.GL_ARRAY_BUFFER, positionVBO);
				gl.glBufferSubData(GL.GL_ARRAY_BUFFER, 0, 16L * nbodies, pos4);
				gl.glBindBuffer(GL.GL_ARRAY_BUFFER, velocityVBO);
				gl.glBufferSubData(GL.GL_ARRAY_BUFFER, 0, 16L * nbodies, vel4);
				gl.glBindBuffer(GL.GL_ARRAY_BUFFER, 0);
			}
		}
	}

	@Override
	public int getNumberOfBodies() {
		return nbodies;
	}
}
