/*
 * rzb200.h — C ABI of the B200-native render path that replaces RayZath's CUDA engine.
 *
 * Drop-in boundary: everything `RayZath::Cuda::Engine::renderWorld`
 * (/root/reference/RayZath/cuda_engine.cuh:21-40, cuda_engine_core.cu:32-128) does on the device
 * is reachable through the entry points below. The C++ host shim that implements
 * `RayZath::Cuda::Engine` on top of them lives in rayzath_b200/host/ (see INTEGRATION.md); Python
 * binds the same symbols with ctypes (rayzath_b200/capi.py).
 *
 * Rules of the boundary: plain pointers and sizes only; no C++ or torch types; every function
 * returns 0 on success and a non-zero code on failure (rzb_last_error() gives the text); the
 * caller owns every host buffer, the context owns every device buffer unless a function says
 * "device pointer"; calls on one context are not re-entrant. There is no CPU fallback: a context
 * cannot be created without a CUDA device.
 *
 * All structs are plain-old-data with explicit sizes (static-asserted in rzb_api.cu and mirrored by
 * numpy dtypes in rayzath_b200/capi.py).
 */
#ifndef RZB200_H
#define RZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RZB_ABI_VERSION 3u /* 2: rzb_scene::flags, own trees, async resolve, incremental update, work counters grew
                              3: sliced multi-process resolve, mesh-tree refit, tree validation (depth / acyclic) */
#define RZB_NO_INDEX 0xFFFFFFFFu
#define RZB_MAX_MATERIALS_PER_INSTANCE 64u /* Instance::materialCapacity(), instance.hpp */

/* error codes */
enum
{
	RZB_OK = 0,
	RZB_ERR_INVALID = 1, /* bad argument / inconsistent scene */
	RZB_ERR_CUDA = 2,    /* a CUDA runtime call failed */
	RZB_ERR_STATE = 3,   /* call order (e.g. render before set_scene) */
	RZB_ERR_NOMEM = 4
};

typedef struct rzb_ctx rzb_ctx;

/* BVH node, 32 B. Same encoding as the reference's device TreeNode
 * (cuda_bvh_tree_node.cuh:11-26) without its 16-byte alignment padding:
 * leaf  <=> count != 0, objects [begin, begin+count);
 * inner <=> count == 0, children at nodes[begin] and nodes[begin+1], split type in bits 30..31
 * (Z=0, Y=1, X=2, Size=3; bvh_tree_node.hpp:21-27). Indices are relative to the owning tree. */
typedef struct rzb_node
{
	float bb_min[3];
	float bb_max[3];
	uint32_t begin;
	uint32_t type_count; /* (split_type << 30) | count */
} rzb_node;

/* Triangle in BVH (leaf-encounter) order, 112 B; what Mesh::reconstruct emits per triangle
 * (cuda_instance.cu:101-147): vertices, per-vertex normals (face normal when the mesh has none),
 * face normal, texture coordinates ((0,0),(0,1),(1,0) when missing), material slot & 0x3F. */
typedef struct rzb_triangle
{
	float v[3][3];
	float n[3][3];
	float face_normal[3];
	float uv[3][2];
	uint32_t material_slot;
} rzb_triangle;

/* One mesh = a node range and a triangle range inside the scene-wide arrays. */
typedef struct rzb_mesh
{
	uint32_t node_offset, node_count;
	uint32_t tri_offset, tri_count;
} rzb_mesh;

/* Instance in top-level BVH order, 100 B (cuda_instance.cuh:167-177): transformation
 * (position, scale, coordinate-system axes = transformationInGroup()), world-space bounding box,
 * mesh id, material table slice, host container index. */
typedef struct rzb_instance
{
	float position[3];
	float scale[3];
	float axis_x[3], axis_y[3], axis_z[3];
	float bb_min[3], bb_max[3];
	uint32_t mesh;            /* index into meshes, RZB_NO_INDEX = no mesh */
	uint32_t material_offset; /* first entry in instance_materials */
	uint32_t material_count;  /* entries (<= 64); slot >= count resolves to the default material */
	uint32_t host_index;      /* idx() in the host Instance container */
} rzb_instance;

/* Uber-material, 64 B (cuda_material.cuh:9-25). color is RGBA in [0,1] (u8/255). Map ids index
 * rzb_scene.maps, RZB_NO_INDEX = no map. */
typedef struct rzb_material
{
	float color[4];
	float metalness, roughness, emission, ior, scattering;
	uint32_t texture, normal_map, metalness_map, roughness_map, emission_map;
	uint32_t _pad[2];
} rzb_material;

enum { RZB_MAP_RGBA8 = 0, RZB_MAP_R8 = 1, RZB_MAP_R32F = 2 };
enum { RZB_FILTER_POINT = 0, RZB_FILTER_LINEAR = 1 };
enum { RZB_ADDRESS_WRAP = 0, RZB_ADDRESS_CLAMP = 1, RZB_ADDRESS_MIRROR = 2, RZB_ADDRESS_BORDER = 3 };

/* Texture / normal / metalness / roughness / emission map (render_parts.hpp:100-225,
 * cuda_buffer.cuh:297-448). pixels is a HOST pointer to width*height texels, row-major, row 0 first. */
typedef struct rzb_map
{
	uint32_t format; /* RZB_MAP_* */
	uint32_t width, height;
	uint32_t filter;  /* RZB_FILTER_* */
	uint32_t address; /* RZB_ADDRESS_* */
	float scale[2];
	float rotation;
	float translation[2];
	uint32_t _pad;
	const void* pixels;
} rzb_map;

/* cuda_direct_light.cuh:14-22 */
typedef struct rzb_direct_light
{
	float direction[3];
	float angular_size;
	float color[3]; /* RGB in [0,1] */
	float emission;
} rzb_direct_light;

/* cuda_spot_light.cuh:15-24 */
typedef struct rzb_spot_light
{
	float position[3];
	float size;
	float direction[3];
	float beam_angle;
	float color[3];
	float emission;
} rzb_spot_light;

/* The whole world, flattened (what World::reconstructResources/Objects mirror to the device,
 * cuda_world.cu). All pointers are HOST pointers; rzb_set_scene copies everything. */
typedef struct rzb_scene
{
	const rzb_node* mesh_nodes;        uint32_t mesh_node_count;
	const rzb_triangle* triangles;     uint32_t triangle_count;
	const uint32_t* tri_host_index;    /* per BVH-order triangle: index in the host mesh's triangle container; may be NULL */
	const rzb_mesh* meshes;            uint32_t mesh_count;
	const rzb_node* instance_nodes;    uint32_t instance_node_count;
	const rzb_instance* instances;     uint32_t instance_count;
	const uint32_t* instance_materials; uint32_t instance_material_count; /* material ids */
	const rzb_material* materials;     uint32_t material_count;
	const rzb_map* maps;               uint32_t map_count;
	const rzb_direct_light* direct_lights; uint32_t direct_light_count;
	const rzb_spot_light* spot_lights; uint32_t spot_light_count;
	rzb_material world_material;       /* World::material(): medium rays start in + sky */
	uint32_t default_material;         /* index into materials used for empty slots */
	uint32_t flags;                    /* RZB_SCENE_* */
} rzb_scene;

enum
{
	RZB_SCENE_REFERENCE_TREES = 0, /* the trees are the reference's own (rzb_build_mesh_bvh or the host World's): every box
	                                  test decides exactly as the reference's arithmetic does -- the parity mode */
	RZB_SCENE_KEEP_GEOMETRY = 2,   /* incremental update: meshes, mesh_nodes, triangles and tri_host_index
	                                  are unchanged since the last full rzb_set_scene on this context -- their fields are
	                                  ignored (may be NULL / 0) and the device copies are reused; instances, the instance tree,
	                                  materials, maps and lights are replaced. Instances may only reference the meshes of that
	                                  upload. Fails with RZB_ERR_STATE when there is no previous full upload or the instance
	                                  tree outgrew the space reserved for it (then upload the whole scene). */
	RZB_SCENE_WIDE_TREES = 4,      /* with RZB_SCENE_OWN_TREES (wide collapse): rzb_set_scene collapses the
	                                  uploaded binary mesh trees into 4-ary ones (a node's children are replaced by its
	                                  grandchildren, largest box first) and the own-tree kernels walk those: one 112-byte fetch and
	                                  four conservative box tests per step, nearest hit first. Same hit records as the binary
	                                  trees except exact ties. At most 2^25 triangles and 15 triangles per leaf. */
	RZB_SCENE_OWN_TREES = 1        /* the mesh trees come from another builder (rzb_build_mesh_bvh_sah): nothing has to
	                                  follow the reference's box decisions, so the kernels use a cheaper conservative
	                                  box test; closest-hit records still equal the reference's except on exact ties */
};

/* Camera (cuda_camera.cuh:112-200, camera.hpp). Axes are the camera coordinate system. */
typedef struct rzb_camera
{
	uint32_t width, height;
	float position[3];
	float axis_x[3], axis_y[3], axis_z[3];
	float fov;              /* radians, full horizontal angle */
	float near_far[2];
	float focal_distance;
	float aperture;
	float exposure_time;
	float temporal_blend;
	uint32_t raycast_pixel[2];
} rzb_camera;

/* RenderConfig (engine_parts.hpp:76-128) + the seed the reference never exposes. */
typedef struct rzb_config
{
	uint32_t spot_light_samples;   /* clamped to >= 1 like cuda_kernel_data.cu:27-28 */
	uint32_t direct_light_samples; /* clamped to >= 1 */
	uint32_t max_depth;            /* 1..255 */
	uint32_t flags;                /* RZB_FLAG_* */
	uint64_t seed;
} rzb_config;

enum
{
	RZB_FLAG_NONE = 0,
	RZB_FLAG_CPU_SEMANTICS = 1, /* follow cpu_engine_kernel.cpp where it differs from the CUDA kernel:
	                               no medium scattering / Beer-Lambert, opaque shadows, texture replaces colour */
	RZB_FLAG_COUNT_WORK = 2,    /* rzb_render counts box tests / triangle tests / shadow rays (rzb_work_counters);
	                               measurement aid, slower kernels */
	RZB_FLAG_SERIAL_STAGES = 8, /* rzb_render launches every kernel of a pass in stream order. Without it (default) the
	                               shadow kernel of pass p runs on a second stream beside the closest-hit kernel of pass
	                               p + 1 (it only adds to the accumulator, which nothing reads before the next shading
	                               kernel; about 2 % faster). With it the per-stage times of rzb_render_stats are exclusive
	                               kernel times: measurement aid */
	RZB_FLAG_TEMPORAL_REPROJECTION = 4 /* Camera::reproject + spacialReprojection (cuda_camera.cuh:390-426,
	                               cuda_postprocess_kernel.cu:5-16): when accumulation restarts (rzb_reset after at least one
	                               rendered pass at the same resolution), the first pass projects every pixel's hit point
	                               into the camera of the frame that is being replaced and, where that frame's depth agrees
	                               within 1 %, adds its accumulator value (rgb sum and sample count) times
	                               rzb_camera::temporal_blend. The reference always does this; here it is opt-in so that
	                               "reset" alone means a clean restart (the C++ drop-in switches it on). The first frame has
	                               no history (the reference reads uninitialised memory there). */
};

/* Closest-hit record (TraversalResult, cuda_render_parts.cuh:946-952), 24 B. */
typedef struct rzb_hit
{
	uint32_t instance; /* host index of the hit instance, RZB_NO_INDEX on miss */
	uint32_t triangle; /* index in the host mesh's triangle container (BVH order if no map was given) */
	float t;           /* ray.near_far.y after traversal */
	float b1, b2;      /* barycentrics */
	uint32_t external; /* det > 0 */
} rzb_hit;

/* Work counters of one trace call (for the algorithmic-bytes figure). */
typedef struct rzb_trace_stats
{
	uint64_t rays;
	uint64_t top_nodes, instances_entered, mesh_nodes, triangles;
} rzb_trace_stats;

typedef struct rzb_render_stats
{
	uint64_t passes;           /* since last reset */
	uint64_t ray_count;        /* passes * width * height  (cuda_render_kernel.cu:122-129) */
	uint64_t shadow_rays;      /* any-hit queries issued by the last pass */
	uint64_t kernel_launches;  /* kernels launched by this context since creation */
	float last_render_ms;      /* device time of the last rzb_render call (CUDA events on the context stream) */
	/* per-stage device time of the last rzb_render call, averaged over its sampled passes (ms per launch) */
	float last_trace_ms, last_shade_ms, last_shadow_ms;
	float last_sort_ms;        /* ray-order pass between shade and the next trace (0 when ray sorting is off) */
	float last_exchange_ms;    /* device time of the last rzb_resolve_sliced kernel (includes waiting for the slowest rank) */
} rzb_render_stats;

/* Work done by rzb_render since the last rzb_reset while RZB_FLAG_COUNT_WORK was set. */
typedef struct rzb_work_counters
{
	uint64_t closest_top_nodes, closest_instances, closest_mesh_nodes, closest_triangles;
	uint64_t shadow_top_nodes, shadow_instances, shadow_mesh_nodes, shadow_triangles;
	uint64_t shadow_rays;
	uint64_t segments; /* closest-hit queries = passes * pixels */
	uint64_t invalid_rays; /* path segments whose sample was non-finite: dropped, path ended (see k_shade) */
	/* SIMT lane utilisation of the whole-warp ray batches: lane_work = sum over rays of (pair steps + triangle tests),
	 * batch_work = sum over 32-ray batches of 32 x the largest such figure in the batch; lane_work / batch_work is the
	 * fraction of lane-time a batch keeps busy if every step cost the same. */
	uint64_t closest_lane_work, closest_batch_work;
	uint64_t shadow_lane_work, shadow_batch_work;
} rzb_work_counters;

/* ---- context ---- */
int rzb_abi_version(void);
/* device: CUDA ordinal. Fails (RZB_ERR_CUDA) when no device is present. */
int rzb_create(int device, rzb_ctx** out);
void rzb_destroy(rzb_ctx* ctx);
/* text of the last error on this context (ctx may be NULL for creation errors) */
const char* rzb_last_error(const rzb_ctx* ctx);
/* Run every kernel and copy of this context on the caller's CUDA stream (a cudaStream_t, e.g. the host
 * framework's current stream; NULL is the legacy default stream, as everywhere in CUDA) when use_caller_stream
 * is non-zero; zero restores the context's private non-blocking stream. Replaces the reference's fixed
 * m_render_stream / m_mirror_stream pair (cuda_engine_core.cu:245-249). */
int rzb_set_stream(rzb_ctx* ctx, void* cuda_stream, int use_caller_stream);

/* ---- world mirror: replaces World::reconstructAll (cuda_world.cu) ---- */
int rzb_set_scene(rzb_ctx* ctx, const rzb_scene* scene);
int rzb_set_camera(rzb_ctx* ctx, const rzb_camera* camera);
int rzb_set_config(rzb_ctx* ctx, const rzb_config* config);
/* Tile split for multi-GPU frames: this context renders only image rows [row_begin, row_end) of the current camera
 * (default: all rows; reset by rzb_set_camera with a new resolution). Rows outside the band stay zero in the
 * accumulator, so the accumulators of disjoint bands (and of sample streams) combine by plain summation. */
int rzb_set_rows(rzb_ctx* ctx, uint32_t row_begin, uint32_t row_end);
/* Interleaved variant (load-balanced: sky rows and geometry rows are spread over all contexts): this context renders
 * the 16-pixel-high chunk rows r with r % count == index. Combines with rzb_set_rows; default index 0, count 1. */
int rzb_set_row_interleave(rzb_ctx* ctx, uint32_t index, uint32_t count);

/* ---- frame: replaces Renderer::renderFunction (cuda_engine_renderer.cu:73-262) ---- */
/* drop accumulated samples, regenerate pixel-centre camera rays (passReset + generateCameraRay). */
int rzb_reset(rzb_ctx* ctx);
/* trace `passes` path segments per pixel (renderFirstPass / renderCumulativePass loop). Asynchronous
 * with respect to the host; every read-back function synchronises. */
int rzb_render(rzb_ctx* ctx, uint32_t passes);
/* tone-map the accumulator (cuda_postprocess_kernel.cu:38-93) and copy results to HOST buffers:
 * rgba8 = width*height*4 bytes, depth = width*height floats (either may be NULL). */
int rzb_resolve(rzb_ctx* ctx, uint8_t* rgba8, float* depth, uint64_t* ray_count);
/* raw float accumulator (rgb sum, alpha = completed paths), width*height*4 floats, HOST buffer. */
int rzb_read_accum(rzb_ctx* ctx, float* rgba_f32);
/* Pipelined hosts (the reference's sync == false: the caller gets the PREVIOUS frame while this one renders,
 * cuda_engine_core.cu:111-121). rzb_resolve_async enqueues the tone map, the copies into the caller's PINNED buffers
 * (rzb_host_alloc; either may be NULL) and the ray-cast pick behind the rendering already enqueued on the context's
 * stream and returns at once; `slot` (0 or 1) names the completion event, so a host can have one frame in flight while
 * it reads the other. ray_count is known on the host and returned immediately. rzb_resolve_wait blocks until that
 * slot's resolve has finished (later work keeps running) and hands back its pick (RZB_NO_INDEX = nothing hit). */
int rzb_resolve_async(rzb_ctx* ctx, uint32_t slot, uint8_t* rgba8_pinned, float* depth_pinned, uint64_t* ray_count);
int rzb_resolve_wait(rzb_ctx* ctx, uint32_t slot, uint32_t* instance, uint32_t* material_slot);
/* page-locked host memory for the asynchronous copies */
int rzb_host_alloc(size_t bytes, void** out);
int rzb_host_free(void* p);
/* mean of the accumulator's alpha channel over the pixels this context renders = completed paths per pixel ("spp",
 * cuda_render_kernel.cu:45,104): what a host that renders "to N spp" stops at. Synchronises. */
int rzb_mean_samples(rzb_ctx* ctx, double* mean_out);
/* device pointer of the linear float4 accumulator (for NCCL reduction by the caller) and its size. */
int rzb_accum_device_ptr(rzb_ctx* ctx, void** device_ptr, size_t* bytes);
/* add `count` float4 pixels from a DEVICE buffer into the accumulator (after a reduce-scatter or P2P read). */
int rzb_accum_add_device(rzb_ctx* ctx, const void* device_rgba_f32, size_t pixel_count);
/* fused multi-GPU resolve: sums the accumulators of `n_peers` other contexts of THIS process over
 * NVLink peer loads while tone-mapping on ctx's device; results as rzb_resolve. */
int rzb_resolve_peers(rzb_ctx* ctx, rzb_ctx* const* peers, uint32_t n_peers,
	uint8_t* rgba8, float* depth, uint64_t* ray_count);
/* One-process-per-GPU variant of the fused resolve: export this context's accumulator as a CUDA IPC handle
 * (64 bytes) ... */
int rzb_accum_ipc_handle(rzb_ctx* ctx, void* handle_out_64_bytes);
/* ... and, on the root rank, open the other ranks' handles and sum + tone-map over NVLink peer loads in ONE
 * kernel. handles = n_peers * 64 bytes. The peers must have finished rendering (barrier) before the call. */
int rzb_resolve_ipc(rzb_ctx* ctx, const void* handles, uint32_t n_peers,
	uint8_t* rgba8, float* depth);
/* Sliced resolve for one process per GPU (no NCCL call, no host round trip; the exchange step of the multi-GPU path):
 * every rank calls rzb_resolve_sliced with the same `world` and its own `rank`; ONE kernel per rank (1) tells all peers
 * through flags in peer memory that its passes are done and waits for theirs, (2) sums slice `rank` of ALL ranks'
 * accumulators over NVLink peer loads, tone-maps it and stores the RGBA8 pixels into rank 0's staging image, (3) tells
 * all peers it is done and waits for theirs -- so when it ends the peers may render into their accumulators again.
 * rzb_exchange_ipc_handle exports the context's exchange buffer (flags + staging image; allocated for the current
 * resolution) as a 64-byte CUDA IPC handle; accum_handles / exchange_handles are world * 64 bytes in rank order
 * (entry [rank] is ignored). Asynchronous: rank 0's copies into its PINNED buffers (either may be NULL; depth is rank
 * 0's own first-pass depth) are enqueued behind the kernel; rzb_resolve_sliced_wait blocks until they are done and
 * reports the device time of the exchange kernel (includes waiting for the slowest rank). Fails with RZB_ERR_STATE when
 * a peer never arrived (the kernel gives up after a spin limit instead of hanging the GPU). */
int rzb_exchange_ipc_handle(rzb_ctx* ctx, void* handle_out_64_bytes);
int rzb_resolve_sliced(rzb_ctx* ctx, uint32_t rank, uint32_t world, const void* accum_handles,
	const void* exchange_handles, uint8_t* rgba8_pinned, float* depth_pinned);
int rzb_resolve_sliced_wait(rzb_ctx* ctx, float* exchange_ms_or_null);
/* pick ray (rayCast kernel, cuda_render_kernel.cu:130-144): instance host index + material slot. */
int rzb_raycast(rzb_ctx* ctx, uint32_t* instance, uint32_t* material_slot);
int rzb_synchronize(rzb_ctx* ctx);
int rzb_get_render_stats(rzb_ctx* ctx, rzb_render_stats* out);
int rzb_get_work_counters(rzb_ctx* ctx, rzb_work_counters* out);
/* human-readable per-stage timings (Engine::timingsString) */
int rzb_timings(rzb_ctx* ctx, char* buf, size_t buf_size);

/* ---- traversal entry points used for ID parity and the roofline kernel ---- */
/* Closest hit for n rays given as HOST arrays: origins[n][3], directions[n][3] (already normalised),
 * near_far[n][2]. World::closestObjectIntersection semantics (cuda_world.cuh:80-90). */
int rzb_trace_closest(rzb_ctx* ctx, const float* origins, const float* directions,
	const float* near_far, uint32_t n, rzb_hit* hits_out, rzb_trace_stats* stats_or_null);
/* Same with DEVICE-resident inputs/outputs (rays as float4 {o.xyz, near}, {d.xyz, far}); returns
 * after enqueueing on the context stream; elapsed_ms_or_null forces a sync and reports CUDA-event time. */
int rzb_trace_closest_device(rzb_ctx* ctx, const void* rays_o_near, const void* rays_d_far,
	uint32_t n, void* hits_out_device, float* elapsed_ms_or_null);
/* Same, with the counting kernel: each 32-byte device hit record additionally carries the ray's own work in its
 * words 5 and 6 (pair steps, triangle tests). Measurement aid. */
int rzb_trace_closest_device_counted(rzb_ctx* ctx, const void* rays_o_near, const void* rays_d_far,
	uint32_t n, void* hits_out_device);
/* Shadow query: World::anyIntersection (cuda_world.cuh:101-104); mask_out[n][4] = RGBA shadow mask. */
int rzb_trace_any(rzb_ctx* ctx, const float* origins, const float* directions,
	const float* near_far, uint32_t n, float* mask_out);
/* pixel-centre camera rays of the current camera (Camera::generateSimpleRay, cuda_camera.cuh:303-328)
 * written to HOST arrays; used to build the fixed primary-ray set. */
int rzb_generate_camera_rays(rzb_ctx* ctx, float* origins, float* directions, float* near_far);

/* ---- host utilities (no device needed) ---- */
/* Triangle BVH exactly as the reference's host builder produces it and Mesh::reconstruct flattens it
 * (component_container.hpp:259-363, cuda_instance.cu:161-220). vertices[nv][3], tris[nt][3] vertex ids.
 * Outputs: nodes (capacity >= 2*nt+1), order[nt] = host triangle index per BVH-order slot.
 * Returns node count through node_count_out. */
int rzb_build_mesh_bvh(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out);
/* OPTIONAL builder, not the reference's tree: binned surface-area-heuristic splits, leaves of at
 * most max_leaf triangles (8 = the reference's leaf size), depth <= 31, same node / order format, so every kernel runs
 * on it unchanged. Closest-hit records equal the reference tree's except on exact-distance ties. */
int rzb_build_mesh_bvh_sah(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt, uint32_t max_leaf,
	rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out);
/* OPTIONAL GPU builder: linear BVH built on `device` (Morton codes, radix sort, Karras' parallel
 * radix tree, bottom-up boxes, leaves of at most max_leaf triangles), same output format and guarantees as
 * rzb_build_mesh_bvh_sah. For rebuild speed (1M triangles: milliseconds of device time, reported through device_ms_out,
 * may be NULL); the SAH tree traces faster. Inputs whose radix tree is deeper than the traversal stack allows are handed
 * to rzb_build_mesh_bvh_sah. */
int rzb_build_mesh_bvh_lbvh(int device, const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	uint32_t max_leaf, rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out,
	float* device_ms_out);
/* Refit: the vertices of a mesh moved, its triangle list did not -- recompute the boxes of an existing
 * tree (any builder's) bottom-up, keeping its topology and triangle order: leaf box = exact min / max of its triangles'
 * vertices, inner box = union of the children's. The result is a valid tree for the deformed mesh (closest-hit records equal
 * a fresh build's except exact ties) but no longer the tree the reference's builder would make for it: upload it with
 * RZB_SCENE_OWN_TREES. nodes is updated in place; order as returned by the builder. O(triangles) on the host
 * (a full GPU rebuild of 1M triangles with rzb_build_mesh_bvh_lbvh takes ~2.5 ms of device time). */
int rzb_refit_mesh_bvh(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	rzb_node* nodes, uint32_t node_count, const uint32_t* order);
/* Instance BVH (bvh_tree_node.hpp:117-215 + cuda_bvh.cuh:86-111): boxes[n][6] = min xyz, max xyz. */
int rzb_build_instance_bvh(const float* boxes, uint32_t n,
	rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out);

/* Coordinate-system axes (x, y, z as 9 floats) from Euler angles, rounded like the reference's host code:
 * order 0 = X, Y, Z (instances: CoordSystem::applyRotation, render_parts.cpp:51-56),
 * order 1 = Z, X, Y (cameras: CoordSystem::lookAt, render_parts.cpp:57-62). */
int rzb_rotation_axes(const float* rotation_xyz, int order, float* axes_out);
/* World-space bounding box of an instance (Instance::calculateBoundingBox, instance.cpp:118-155):
 * bbox_out = min xyz, max xyz. */
int rzb_instance_bbox(const float* vertices, uint32_t nv, const float* position, const float* scale,
	const float* axes, float* bbox_out);
/* Face normals (Triangle::calculateNormal, mesh_component.cpp:19-26): normals_out[nt][3]. */
int rzb_face_normals(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt, float* normals_out);

#ifdef __cplusplus
}
#endif
#endif /* RZB200_H */
