/*
 * JNI shim for JDK 8-era hosts (the reference targets Java 8, .settings/org.eclipse.jdt.core.prefs:3-11): one
 * trivial forwarding function per C-ABI entry of include/bhstep.h that ch.fhnw.woipv.nbody.simulation.gpu.BhStepJni
 * declares `native`.  NOT COMPILED in this repository's build image (no JDK, hence no jni.h); build where a JDK is
 * installed:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I include java/jni/bhstep_jni.c \
 *       -L gpu_nbody_b200 -lbhstep -o libbhstep_jni.so
 * Array lengths are validated before anything is touched; the data crosses through Get/SetXxxArrayRegion into malloc'd
 * staging, so no JNI critical region is ever held across a CUDA call (cudaMalloc, pageable copies, stream syncs).
 */
#include <stdlib.h>
#include <jni.h>
#include <stdint.h>

#include "bhstep.h"

#define J(name) Java_ch_fhnw_woipv_nbody_simulation_gpu_BhStepJni_##name

JNIEXPORT jlong JNICALL J(create)(JNIEnv *env, jclass c, jint nbodies, jfloat theta, jfloat eps2, jfloat dt, jint vote, jint device) {
    bh_sim *sim = NULL;
    (void)env; (void)c;
    return bh_create(&sim, nbodies, theta, eps2, dt, vote, device) == 0 ? (jlong)(intptr_t)sim : 0; /* 0: see lastError(0) */
}

JNIEXPORT void JNICALL J(destroy)(JNIEnv *env, jclass c, jlong sim) { (void)env; (void)c; bh_destroy((bh_sim *)(intptr_t)sim); }

JNIEXPORT jstring JNICALL J(lastError)(JNIEnv *env, jclass c, jlong sim) {
    (void)c;
    return (*env)->NewStringUTF(env, bh_last_error((bh_sim *)(intptr_t)sim));
}

JNIEXPORT jint JNICALL J(numberOfNodes)(JNIEnv *env, jclass c, jint nbodies) { (void)env; (void)c; return bh_number_of_nodes(nbodies); }

/* loadBuffers, GPUBH:155-170 */
JNIEXPORT jint JNICALL J(upload)(JNIEnv *env, jclass c, jlong sim, jfloatArray x, jfloatArray y, jfloatArray z, jfloatArray vx,
                                 jfloatArray vy, jfloatArray vz, jfloatArray mass) {
    jfloatArray arr[7] = {x, y, z, vx, vy, vz, mass};
    const int n = bh_number_of_bodies((bh_sim *)(intptr_t)sim);
    float *stage;
    int i, rc;
    (void)c;
    if (n <= 0) return BH_ERR_ARG;
    for (i = 0; i < 7; ++i)
        if (!arr[i] || (*env)->GetArrayLength(env, arr[i]) < n) return BH_ERR_ARG;
    stage = (float *)malloc(sizeof(float) * 7 * (size_t)n);
    if (!stage) return BH_ERR_ALLOC;
    for (i = 0; i < 7; ++i) (*env)->GetFloatArrayRegion(env, arr[i], 0, n, stage + (size_t)i * n);
    rc = bh_upload((bh_sim *)(intptr_t)sim, stage, stage + (size_t)n, stage + 2 * (size_t)n, stage + 3 * (size_t)n, stage + 4 * (size_t)n,
                   stage + 5 * (size_t)n, stage + 6 * (size_t)n);
    free(stage);
    return rc;
}

/* step(), GPUBH:249-271 */
JNIEXPORT jint JNICALL J(step)(JNIEnv *env, jclass c, jlong sim, jint nsteps) { (void)env; (void)c; return bh_step((bh_sim *)(intptr_t)sim, nsteps); }

/* executeSimulationKernel, GPUBH:258-263; stage 0..5 */
JNIEXPORT jint JNICALL J(stage)(JNIEnv *env, jclass c, jlong sim, jint stage) {
    bh_sim *s = (bh_sim *)(intptr_t)sim;
    (void)env; (void)c;
    switch (stage) {
    case 0: return bh_bounding_box(s);
    case 1: return bh_build_tree(s);
    case 2: return bh_summarize(s);
    case 3: return bh_sort(s);
    case 4: return bh_calculate_force(s);
    case 5: return bh_integrate(s);
    default: return BH_ERR_ARG;
    }
}

/* readBuffer + getData, GPUBH:277-278,294-295,306-312 */
JNIEXPORT jint JNICALL J(readFloats)(JNIEnv *env, jclass c, jlong sim, jint which, jfloatArray dst, jint count) {
    float *stage;
    int rc;
    (void)c;
    if (!dst || count < 0 || count > (*env)->GetArrayLength(env, dst)) return BH_ERR_ARG;
    if (count == 0) return 0;
    stage = (float *)malloc(sizeof(float) * (size_t)count);
    if (!stage) return BH_ERR_ALLOC;
    rc = bh_read((bh_sim *)(intptr_t)sim, which, stage, count);
    if (rc == 0) (*env)->SetFloatArrayRegion(env, dst, 0, count, stage);
    free(stage);
    return rc;
}

JNIEXPORT jint JNICALL J(readInts)(JNIEnv *env, jclass c, jlong sim, jint which, jintArray dst, jint count) {
    jint *stage;
    int rc;
    (void)c;
    if (!dst || count < 0 || count > (*env)->GetArrayLength(env, dst)) return BH_ERR_ARG;
    if (count == 0) return 0;
    stage = (jint *)malloc(sizeof(jint) * (size_t)count);
    if (!stage) return BH_ERR_ALLOC;
    rc = bh_read((bh_sim *)(intptr_t)sim, which, stage, count);
    if (rc == 0) (*env)->SetIntArrayRegion(env, dst, 0, count, stage);
    free(stage);
    return rc;
}

/* copyVertices into direct ByteBuffers (e.g. mapped GL buffers), GPUBH:265-266 */
JNIEXPORT jint JNICALL J(copyVertices)(JNIEnv *env, jclass c, jlong sim, jobject pos4, jobject vel4) {
    const jlong need = 16 * (jlong)bh_number_of_bodies((bh_sim *)(intptr_t)sim);
    (void)c;
    if ((pos4 && (!(*env)->GetDirectBufferAddress(env, pos4) || (*env)->GetDirectBufferCapacity(env, pos4) < need)) ||
        (vel4 && (!(*env)->GetDirectBufferAddress(env, vel4) || (*env)->GetDirectBufferCapacity(env, vel4) < need)))
        return BH_ERR_ARG; /* not direct, or too small for nbodies float4 */
    return bh_copy_vertices((bh_sim *)(intptr_t)sim, pos4 ? (float *)(*env)->GetDirectBufferAddress(env, pos4) : NULL,
                            vel4 ? (float *)(*env)->GetDirectBufferAddress(env, vel4) : NULL);
}
