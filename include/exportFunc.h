/*
 * exportFunc.h -- the Unity-facing C ABI of DragPoserDLL, re-implemented natively on the
 * B200 engine.  Symbol names, argument order and struct layouts are those of the reference
 * (DragPoserDLL/exportFunc.h:61-70, structs DragPoserDLL/utils.h:13-41), so the Unity side
 * (DragPoserUnity/Assets/Scripts/Core/DragPoserDLL.cs:10-29 [DllImport]) binds unchanged.
 * The reference shim embeds CPython and forwards every call to python/src/run_drag.py; this
 * one parses the BVH header and the model file natively and drives the CUDA engine through
 * include/dp_engine.h: a frame is staging copy -> one launch -> one sync, no interpreter.
 *
 * Semantics (python/src/run_drag.py:30-176): targets are positions relative to the character
 * root of the previous frame and world rotations (w,x,y,z quaternions), BVH right-handed
 * coordinates; resultPose receives the J parent-local quaternions, resultGlobalPos the root.
 * Every buffer is owned by the caller; nothing is retained across calls.  No function throws
 * or aborts across the boundary: failures are logged to stderr (and to the file named by
 * DRAGPOSER_LOG if set) and leave the outputs untouched; dp_last_status() reports them.
 */
#ifndef DRAGPOSER_EXPORTFUNC_H
#define DRAGPOSER_EXPORTFUNC_H

#if _WIN32
#define DP_EXPORT __declspec(dllexport)
#else
#define DP_EXPORT __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct quaternion { float w, x, y, z; } quaternion;   /* utils.h:13-22 */
typedef struct float3 { float x, y, z; } float3;              /* utils.h:24-32 */
typedef struct float2 { float x, y; } float2;                 /* utils.h:34-41 */
typedef struct DragPoser DragPoser;                           /* opaque session (exportFunc.h:11-59) */

DP_EXPORT DragPoser* init_drag_poser(void);                                            /* exportFunc.cpp:5-20 */
DP_EXPORT void set_reference_skeleton(DragPoser* dragPoser, char* bvhPath);            /* :22-28 */
DP_EXPORT void load_models(DragPoser* dragPoser, char* modelPath);                     /* :30-34 */
DP_EXPORT void set_mask_and_weights(DragPoser* dragPoser, float* mask, float2* weights); /* :36-45 */
DP_EXPORT void init_drag_model(DragPoser* dragPoser, float3 initialGlobalPos, quaternion initialGlobalRot); /* :47-53 */
DP_EXPORT void set_optim_params(DragPoser* dragPoser, float stopEpsPos, float stopEpsRot, int maxIter, float lr); /* :55-59 */
DP_EXPORT void set_lambdas(DragPoser* dragPoser, float lambdaRot, float lambdaTemporal, int temporalFutureWindow); /* :61-65 */
DP_EXPORT void set_global_pos(DragPoser* dragPoser, float3 globalPos);                 /* :67-72 */
DP_EXPORT void drag_pose(DragPoser* dragPoser, int nEndEffectors, float3* targetEEPos, quaternion* targetEERot,
                         quaternion* resultPose, float3* resultGlobalPos);             /* :74-95 */
DP_EXPORT void destroy_drag_poser(DragPoser* dragPoser);                               /* :97-100 */

/* Additions (not part of the reference ABI; safe to ignore from Unity). */
DP_EXPORT int dp_last_status(const DragPoser* dragPoser);            /* 0 = last call succeeded */
DP_EXPORT const char* dp_last_message(const DragPoser* dragPoser);
DP_EXPORT int dp_get_num_joints(const DragPoser* dragPoser);
DP_EXPORT int dp_get_num_endeffectors(const DragPoser* dragPoser);
DP_EXPORT void dp_set_initial_latent(DragPoser* dragPoser, const float* latent24);   /* overrides the encoder draw of init_drag_model */
DP_EXPORT void dp_get_initial_latent(const DragPoser* dragPoser, float* latent24);
DP_EXPORT int dp_get_last_iterations(DragPoser* dragPoser);          /* optimisation iterations the last drag_pose ran (-1 on error) */
DP_EXPORT void dp_set_min_loss_increment(DragPoser* dragPoser, double minLossIncr);  /* DragPose.run's min_loss_incr (run_drag.py keeps the
                                                                         default 1e-5 and the reference ABI has no setter); -inf disables */

#ifdef __cplusplus
}
#endif
#endif
