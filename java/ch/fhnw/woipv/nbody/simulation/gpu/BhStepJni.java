package ch.fhnw.woipv.nbody.simulation.gpu;

import java.nio.ByteBuffer;

/**
 * JNI flavour of the binding for Java 8 hosts (java/jni/bhstep_jni.c); same calls as {@link BhStep} (Panama FFM).
 * NOT COMPILED in this repository's build image (no JDK).
 */
public final class BhStepJni {
	static {
		System.loadLibrary("bhstep_jni");
	}

	private BhStepJni() {
	}

	public static native long create(int nbodies, float theta, float eps2, float dt, int voteWidth, int device);

	public static native void destroy(long sim);

	public static native String lastError(long sim);

	public static native int numberOfNodes(int nbodies);

	public static native int upload(long sim, float[] x, float[] y, float[] z, float[] vx, float[] vy, float[] vz, float[] mass);

	public static native int step(long sim, int nsteps);

	public static native int stage(long sim, int stage);

	public static native int readFloats(long sim, int which, float[] dst, int count);

	public static native int readInts(long sim, int which, int[] dst, int count);

	public static native int copyVertices(long sim, ByteBuffer pos4, ByteBuffer vel4);
}
