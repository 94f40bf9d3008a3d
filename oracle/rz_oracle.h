/*
 * rz_oracle.h -- CPU restatement (plain C) of the reference's algorithm for the hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (rayzath_b200/) links, imports or executes this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may. It exists so
 * that the CUDA path can be checked on a box where /root/reference is absent, and it is itself pinned
 * against the reference's own CPU engine compiled from its sources (oracle/_ref, see oracle/Makefile) by
 * tests/test_oracle.py and the committed vectors in tests/golden/.
 *
 * PARITY PIN: the reference ships no golden vectors for this path; the pin is "outputs of the
 * reference itself run here" (oracle/_ref/rz_ref_tool trace / traceany / render). One level remains
 * unpinned: the un-vendored Math library's Normalize/Magnitude (restated in oracle/shim/vec3.h).
 *
 * The scene is the flattened C-ABI form (include/rzb200.h): the reference's own trees in the reference's
 * own flattening order, so traversal order == the reference's.
 */
#ifndef RZ_ORACLE_H
#define RZ_ORACLE_H

#include "../include/rzb200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum
{
	RZO_ORDER_CPU = 0, /* children visited first -> second (cpu_engine_kernel.cpp:254-277, 333-349) */
	RZO_ORDER_CUDA = 1 /* near child first by ray sign on the split axis (cuda_instance.cuh:49-65, cuda_bvh.cuh:129-145) */
};
enum
{
	RZO_MINMAX_SELECT = 0, /* a < b ? a : b  (CPU engine, render_parts.cpp:206-211) */
	RZO_MINMAX_FMINF = 1   /* fminf / fmaxf   (CUDA engine, cuda_render_parts.cuh:1178-1191) */
};

/* Closest hit per ray. origins[n][3], directions[n][3], near_far[n][2]; hits_out[n].
 * stats_or_null accumulates box tests / triangle tests of THIS traversal order (for algorithmic bytes). */
void rzo_trace_closest(const rzb_scene* scene, const float* origins, const float* directions, const float* near_far,
	uint32_t n, int order, int minmax, rzb_hit* hits_out, rzb_trace_stats* stats_or_null);

/* Shadow query with the CPU engine's semantics: any intersected triangle makes the mask 0
 * (cpu_engine_kernel.cpp:398-481). masks_out[n][4]. */
void rzo_trace_any(const rzb_scene* scene, const float* origins, const float* directions, const float* near_far,
	uint32_t n, int minmax, float* masks_out);

/* Pixel-centre camera rays (Kernel::generateSimpleRay, cpu_engine_kernel.cpp:180-204). Arrays of width*height. */
void rzo_camera_rays(const rzb_camera* camera, float* origins, float* directions, float* near_far);

/* Tone map (cpu_engine_renderer.cpp:224-235 / cuda_postprocess_kernel.cu:38-58): accum[n][4] -> rgba8[n][4]. */
void rzo_tonemap(const float* accum, uint32_t n, float aperture, float exposure_time, uint8_t* rgba8);

/* number of worker threads the trace functions use (online CPUs, or the RZO_THREADS environment variable) */
int rzo_threads(void);

#ifdef __cplusplus
}
#endif
#endif
