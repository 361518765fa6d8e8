package ch.fhnw.woipv.nbody.simulation.gpu;

import com.jogamp.opengl.GL3;

import ch.fhnw.woipv.nbody.simulation.AbstractNBodySimulation;
import ch.fhnw.woipv.nbody.simulation.universe.UniverseGenerator;

/**
 * Drop-in for GPUBarnesHutNBodySimulation: same constructor and the four
 * AbstractNBodySimulation methods, the OpenCL calls replaced by libbhstep.so.
 * NOT COMPILED in this repository's build image (no JDK, no JOGL jars on the class path there).
 * In NBodyVisualizer.java:208-221 replace `new GPUBarnesHutNBodySimulation(mode, n, generator)` by
 * `new GPUBarnesHutNBodySimulationB200(mode, n, generator)`.
 */
public class GPUBarnesHutNBodySimulationB200 extends AbstractNBodySimulation {
	private final int nbodies;
	private final UniverseGenerator universeGenerator;
	private BhStep bh;
	private java.lang.foreign.MemorySegment pos4, vel4;
	private final java.lang.foreign.Arena arena = java.lang.foreign.Arena.ofShared();

	public GPUBarnesHutNBodySimulationB200(final Mode mode, final int nbodies, final UniverseGenerator generator) {
		super(mode);
		this.nbodies = nbodies;
		this.universeGenerator = generator;
	}

	@Override
	public void init(final GL3 gl) { // GPUBH:111-151
		final int numberOfNodes = BhStep.numberOfNodes(nbodies);
		final float[] x = new float[numberOfNodes + 1], y = new float[numberOfNodes + 1], z = new float[numberOfNodes + 1];
		final float[] vx = new float[numberOfNodes + 1], vy = new float[numberOfNodes + 1], vz = new float[numberOfNodes + 1];
		final float[] mass = new float[numberOfNodes + 1];
		universeGenerator.generate(0, nbodies, x, y, z, vx, vy, vz, mass);
		// theta = 0.5: the reference's commented-out THETA (0.5f * 0.5f); for the shipped THETA (1.5f) call bh_set_theta_macro
		bh = new BhStep(nbodies, 0.5f, 0.0025f, 0.025f, 16, 0);
		bh.upload(x, y, z, vx, vy, vz, mass);
	}

	@Override
	public void initGLBuffers(final GL3 gl, final int positionVBO, final int velocityVBO) { // GPUBH:230-246
		// Host staging; with GL: glMapBuffer the two VBOs and pass the mapped addresses instead
		pos4 = arena.allocate(java.lang.foreign.ValueLayout.JAVA_FLOAT, 4L * nbodies);
		vel4 = arena.allocate(java.lang.foreign.ValueLayout.JAVA_FLOAT, 4L * nbodies);
	}

	@Override
	public void step() { // GPUBH:249-271
		bh.step(1);
		if (mode == Mode.GL_INTEROP)
			bh.copyVertices(pos4, vel4);
	}

	@Override
	public int getNumberOfBodies() {
		return nbodies;
	}
}
