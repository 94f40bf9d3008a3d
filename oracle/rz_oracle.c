/*
 * rz_oracle.c -- CPU restatement (plain C, fp32, no FMA contraction: build with -ffp-contract=off) of the
 * reference's traversal, camera-ray and tone-map arithmetic. TEST INFRASTRUCTURE ONLY: see rz_oracle.h.
 *
 * Every function names the reference lines it follows (paths relative to /root/reference/RayZath).
 */
#include "rz_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- a small pthread parallel-for (this image's gcc has no libgomp): ranges of 256 rays pulled with an atomic ---- */
typedef void (*range_fn)(void* arg, int64_t begin, int64_t end);
typedef struct
{
	range_fn fn;
	void* arg;
	int64_t n;
	atomic_llong next;
} pf_job;
static int pf_thread_count(void)
{
	const char* env = getenv("RZO_THREADS");
	long t = env ? atol(env) : sysconf(_SC_NPROCESSORS_ONLN);
	if (t < 1) t = 1;
	if (t > 256) t = 256;
	return (int)t;
}
static void* pf_worker(void* p)
{
	pf_job* job = (pf_job*)p;
	for (;;)
	{
		const int64_t b = (int64_t)atomic_fetch_add(&job->next, 256);
		if (b >= job->n) break;
		job->fn(job->arg, b, b + 256 < job->n ? b + 256 : job->n);
	}
	return NULL;
}
static void parallel_for(int64_t n, range_fn fn, void* arg)
{
	pf_job job;
	job.fn = fn; job.arg = arg; job.n = n;
	atomic_init(&job.next, 0);
	int t = pf_thread_count();
	if ((int64_t)t * 256 > n) t = (int)((n + 255) / 256);
	if (t <= 1) { pf_worker(&job); return; }
	pthread_t th[256];
	int started = 0;
	for (int i = 0; i < t - 1; ++i)
		if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
	pf_worker(&job);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

typedef struct { float x, y, z; } v3;

static inline v3 v3_sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
/* Math::vec3::DotProduct / CrossProduct as restated in oracle/shim/vec3.h (left to right) */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 v3_cross(v3 a, v3 b)
{
	v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
	return r;
}

typedef struct
{
	v3 o, d;
	float near_, far_;
} ray_t;

typedef struct
{
	const rzb_scene* sc;
	int order, minmax;
	uint64_t top_nodes, instances, mesh_nodes, triangles;
} ctx_t;

static inline float sel_min(float a, float b, int mode) { return mode == RZO_MINMAX_FMINF ? fminf(a, b) : (a < b ? a : b); }
static inline float sel_max(float a, float b, int mode) { return mode == RZO_MINMAX_FMINF ? fmaxf(a, b) : (a > b ? a : b); }

/* BoundingBox::rayIntersection, render_parts.cpp:197-217 (CPU) / cuda_render_parts.cuh:1178-1191 (CUDA) */
static int box_hit(const float* mn, const float* mx, const ray_t* r, int mode)
{
	const float t1 = (mn[0] - r->o.x) / r->d.x;
	const float t2 = (mx[0] - r->o.x) / r->d.x;
	const float t3 = (mn[1] - r->o.y) / r->d.y;
	const float t4 = (mx[1] - r->o.y) / r->d.y;
	const float t5 = (mn[2] - r->o.z) / r->d.z;
	const float t6 = (mx[2] - r->o.z) / r->d.z;
	const float tmin = sel_max(sel_max(sel_min(t1, t2, mode), sel_min(t3, t4, mode), mode), sel_min(t5, t6, mode), mode);
	const float tmax = sel_min(sel_min(sel_max(t1, t2, mode), sel_max(t3, t4, mode), mode), sel_max(t5, t6, mode), mode);
	return !(tmax < r->near_ || tmin > tmax || tmin > r->far_);
}

typedef struct
{
	uint32_t tri; /* scene-wide BVH-order triangle index, RZB_NO_INDEX = none */
	uint32_t inst;
	float b1, b2;
	int external;
} result_t;

/* Triangle::closestIntersection, mesh_component.cpp:52-83 */
static void tri_closest(const rzb_triangle* t, uint32_t index, ray_t* r, result_t* res)
{
	const v3 v1 = {t->v[0][0], t->v[0][1], t->v[0][2]};
	const v3 v2 = {t->v[1][0], t->v[1][1], t->v[1][2]};
	const v3 v3_ = {t->v[2][0], t->v[2][1], t->v[2][2]};
	const v3 e1 = v3_sub(v2, v1), e2 = v3_sub(v3_, v1);
	const v3 pvec = v3_cross(r->d, e2);
	float det = v3_dot(e1, pvec);
	det += (float)((det > -1.0e-7f) & (det < 1.0e-7f)) * 1.0e-7f;
	const float inv_det = 1.0f / det;
	const v3 tvec = v3_sub(r->o, v1);
	const float b1 = v3_dot(tvec, pvec) * inv_det;
	if (b1 < 0.0f || b1 > 1.0f) return;
	const v3 qvec = v3_cross(tvec, e1);
	const float b2 = v3_dot(r->d, qvec) * inv_det;
	if (b2 < 0.0f || b1 + b2 > 1.0f) return;
	const float tt = v3_dot(e2, qvec) * inv_det;
	if (tt <= r->near_ || tt >= r->far_) return;
	r->far_ = tt;
	res->tri = index;
	res->external = det > 0.0f;
	res->b1 = b1;
	res->b2 = b2;
}
/* Triangle::anyIntersection, mesh_component.cpp:84-113 */
static int tri_any(const rzb_triangle* t, const ray_t* r)
{
	ray_t copy = *r;
	result_t res;
	res.tri = RZB_NO_INDEX;
	tri_closest(t, 0, &copy, &res);
	return res.tri != RZB_NO_INDEX;
}

static inline uint32_t node_count(const rzb_node* n) { return n->type_count & 0x3FFFFFFFu; }
static inline uint32_t node_type(const rzb_node* n) { return n->type_count >> 30; }
static inline uint32_t sign_bits(v3 d)
{
	return ((uint32_t)(d.x < 0.0f) << 2) | ((uint32_t)(d.y < 0.0f) << 1) | (uint32_t)(d.z < 0.0f);
}

/* Kernel::closestIntersection(const Mesh&, ...), cpu_engine_kernel.cpp:329-352: every node tests its own box on
 * entry; children first -> second. RZO_ORDER_CUDA visits the near child first (cuda_instance.cuh:35-91). */
static void mesh_closest(ctx_t* c, const rzb_mesh* mesh, uint32_t node_idx, ray_t* r, result_t* res, uint32_t sbits)
{
	const rzb_node* n = c->sc->mesh_nodes + mesh->node_offset + node_idx;
	c->mesh_nodes++;
	if (!box_hit(n->bb_min, n->bb_max, r, c->minmax)) return;
	const uint32_t count = node_count(n);
	if (count)
	{
		for (uint32_t i = 0; i < count; ++i)
		{
			const uint32_t ti = mesh->tri_offset + n->begin + i;
			c->triangles++;
			tri_closest(c->sc->triangles + ti, ti, r, res);
		}
		return;
	}
	const uint32_t flip = c->order == RZO_ORDER_CUDA ? ((sbits >> node_type(n)) & 1u) : 0u;
	mesh_closest(c, mesh, n->begin + flip, r, res, sbits);
	mesh_closest(c, mesh, n->begin + (flip ^ 1u), r, res, sbits);
}

/* Transformation::transformG2L (render_parts.cpp:113-121 -> CoordSystem::transformBackward :40-47) +
 * Kernel::closestIntersection(instance...), cpu_engine_kernel.cpp:297-328 */
static void to_local(const rzb_instance* in, const ray_t* w, ray_t* l, float* length_factor)
{
	const v3 pos = {in->position[0], in->position[1], in->position[2]};
	const v3 p = v3_sub(w->o, pos);
	const float* ax = in->axis_x; const float* ay = in->axis_y; const float* az = in->axis_z;
	v3 o = {ax[0] * p.x + ax[1] * p.y + ax[2] * p.z, ay[0] * p.x + ay[1] * p.y + ay[2] * p.z, az[0] * p.x + az[1] * p.y + az[2] * p.z};
	o.x /= in->scale[0]; o.y /= in->scale[1]; o.z /= in->scale[2];
	v3 d = {ax[0] * w->d.x + ax[1] * w->d.y + ax[2] * w->d.z, ay[0] * w->d.x + ay[1] * w->d.y + ay[2] * w->d.z,
		az[0] * w->d.x + az[1] * w->d.y + az[2] * w->d.z};
	d.x /= in->scale[0]; d.y /= in->scale[1]; d.z /= in->scale[2];
	const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z); /* Magnitude(), shim/vec3.h */
	l->o = o;
	l->near_ = w->near_ * len;
	l->far_ = w->far_ * len;
	d.x /= len; d.y /= len; d.z /= len; /* Normalize() */
	l->d = d;
	*length_factor = len;
}

static void instance_closest(ctx_t* c, uint32_t inst_idx, ray_t* r, result_t* res)
{
	const rzb_instance* in = c->sc->instances + inst_idx;
	c->instances++;
	if (!box_hit(in->bb_min, in->bb_max, r, c->minmax)) return;
	ray_t local;
	float len;
	to_local(in, r, &local, &len);
	if (in->mesh == RZB_NO_INDEX) return;
	const rzb_mesh* mesh = c->sc->meshes + in->mesh;
	result_t lres = *res;
	lres.tri = RZB_NO_INDEX;
	if (mesh->node_count) mesh_closest(c, mesh, 0, &local, &lres, sign_bits(local.d));
	if (lres.tri != RZB_NO_INDEX)
	{
		*res = lres;
		res->inst = inst_idx;
		r->near_ = local.near_ / len;
		r->far_ = local.far_ / len;
	}
}

/* Kernel::traverseWorld, cpu_engine_kernel.cpp:254-277 (children's boxes are tested by the parent) */
static void world_closest(ctx_t* c, uint32_t node_idx, ray_t* r, result_t* res, uint32_t sbits)
{
	const rzb_node* n = c->sc->instance_nodes + node_idx;
	const uint32_t count = node_count(n);
	if (count)
	{
		for (uint32_t i = 0; i < count; ++i) instance_closest(c, n->begin + i, r, res);
		return;
	}
	const uint32_t flip = c->order == RZO_ORDER_CUDA ? ((sbits >> node_type(n)) & 1u) : 0u;
	const rzb_node* a = c->sc->instance_nodes + n->begin + flip;
	c->top_nodes++;
	if (box_hit(a->bb_min, a->bb_max, r, c->minmax)) world_closest(c, n->begin + flip, r, res, sbits);
	const rzb_node* b = c->sc->instance_nodes + n->begin + (flip ^ 1u);
	c->top_nodes++;
	if (box_hit(b->bb_min, b->bb_max, r, c->minmax)) world_closest(c, n->begin + (flip ^ 1u), r, res, sbits);
}

typedef struct
{
	const rzb_scene* scene;
	const float* origins; const float* directions; const float* near_far;
	int order, minmax;
	rzb_hit* hits_out;
	float* masks_out;
	atomic_ullong s_top, s_inst, s_mesh, s_tri;
} trace_job;

static void closest_range(void* arg, int64_t begin, int64_t end)
{
	trace_job* j = (trace_job*)arg;
	const rzb_scene* scene = j->scene;
	uint64_t s_top = 0, s_inst = 0, s_mesh = 0, s_tri = 0;
	for (int64_t i = begin; i < end; ++i)
	{
		ctx_t c;
		memset(&c, 0, sizeof(c));
		c.sc = scene; c.order = j->order; c.minmax = j->minmax;
		ray_t r;
		r.o.x = j->origins[3 * i]; r.o.y = j->origins[3 * i + 1]; r.o.z = j->origins[3 * i + 2];
		r.d.x = j->directions[3 * i]; r.d.y = j->directions[3 * i + 1]; r.d.z = j->directions[3 * i + 2];
		r.near_ = j->near_far[2 * i]; r.far_ = j->near_far[2 * i + 1];
		result_t res;
		res.tri = RZB_NO_INDEX; res.inst = RZB_NO_INDEX; res.b1 = res.b2 = 0.0f; res.external = 0;
		/* Kernel::closestIntersection(RangedRay&, SurfaceProperties&), cpu_engine_kernel.cpp:279-289 */
		if (scene->instance_count != 0 && scene->instance_node_count != 0)
		{
			const rzb_node* root = scene->instance_nodes;
			c.top_nodes++;
			if (box_hit(root->bb_min, root->bb_max, &r, j->minmax)) world_closest(&c, 0, &r, &res, sign_bits(r.d));
		}
		rzb_hit h;
		memset(&h, 0, sizeof(h));
		h.instance = RZB_NO_INDEX; h.triangle = RZB_NO_INDEX;
		h.t = r.far_;
		if (res.inst != RZB_NO_INDEX)
		{
			h.instance = scene->instances[res.inst].host_index;
			h.triangle = scene->tri_host_index ? scene->tri_host_index[res.tri] : res.tri;
			h.b1 = res.b1; h.b2 = res.b2;
			h.external = (uint32_t)res.external;
		}
		j->hits_out[i] = h;
		s_top += c.top_nodes; s_inst += c.instances; s_mesh += c.mesh_nodes; s_tri += c.triangles;
	}
	atomic_fetch_add(&j->s_top, s_top); atomic_fetch_add(&j->s_inst, s_inst);
	atomic_fetch_add(&j->s_mesh, s_mesh); atomic_fetch_add(&j->s_tri, s_tri);
}

void rzo_trace_closest(const rzb_scene* scene, const float* origins, const float* directions, const float* near_far,
	uint32_t n, int order, int minmax, rzb_hit* hits_out, rzb_trace_stats* stats)
{
	trace_job j;
	j.scene = scene; j.origins = origins; j.directions = directions; j.near_far = near_far;
	j.order = order; j.minmax = minmax; j.hits_out = hits_out; j.masks_out = NULL;
	atomic_init(&j.s_top, 0); atomic_init(&j.s_inst, 0); atomic_init(&j.s_mesh, 0); atomic_init(&j.s_tri, 0);
	parallel_for((int64_t)n, closest_range, &j);
	if (stats)
	{
		stats->rays = n;
		stats->top_nodes = atomic_load(&j.s_top); stats->instances_entered = atomic_load(&j.s_inst);
		stats->mesh_nodes = atomic_load(&j.s_mesh); stats->triangles = atomic_load(&j.s_tri);
	}
}

/* Kernel::anyIntersection(const Mesh&, ...), cpu_engine_kernel.cpp:452-481 */
static int mesh_any(const rzb_scene* sc, const rzb_mesh* mesh, uint32_t node_idx, const ray_t* r, int minmax)
{
	const rzb_node* n = sc->mesh_nodes + mesh->node_offset + node_idx;
	if (!box_hit(n->bb_min, n->bb_max, r, minmax)) return 0;
	const uint32_t count = node_count(n);
	if (count)
	{
		for (uint32_t i = 0; i < count; ++i)
			if (tri_any(sc->triangles + mesh->tri_offset + n->begin + i, r)) return 1;
		return 0;
	}
	return mesh_any(sc, mesh, n->begin, r, minmax) || mesh_any(sc, mesh, n->begin + 1, r, minmax);
}
/* Kernel::anyIntersection(const Instance&, ...), cpu_engine_kernel.cpp:436-451 */
static int instance_any(const rzb_scene* sc, uint32_t inst_idx, const ray_t* r, int minmax)
{
	const rzb_instance* in = sc->instances + inst_idx;
	if (!box_hit(in->bb_min, in->bb_max, r, minmax)) return 0;
	ray_t local;
	float len;
	to_local(in, r, &local, &len);
	if (in->mesh == RZB_NO_INDEX) return 0;
	const rzb_mesh* mesh = sc->meshes + in->mesh;
	return mesh->node_count ? mesh_any(sc, mesh, 0, &local, minmax) : 0;
}
/* Kernel::anyIntersection(const RangedRay&), cpu_engine_kernel.cpp:398-435 */
static int world_any(const rzb_scene* sc, uint32_t node_idx, const ray_t* r, int minmax)
{
	const rzb_node* n = sc->instance_nodes + node_idx;
	const uint32_t count = node_count(n);
	if (count)
	{
		for (uint32_t i = 0; i < count; ++i)
			if (instance_any(sc, n->begin + i, r, minmax)) return 1;
		return 0;
	}
	const rzb_node* a = sc->instance_nodes + n->begin;
	if (box_hit(a->bb_min, a->bb_max, r, minmax) && world_any(sc, n->begin, r, minmax)) return 1;
	const rzb_node* b = sc->instance_nodes + n->begin + 1;
	return box_hit(b->bb_min, b->bb_max, r, minmax) && world_any(sc, n->begin + 1, r, minmax);
}

static void any_range(void* arg, int64_t begin, int64_t end)
{
	trace_job* j = (trace_job*)arg;
	const rzb_scene* scene = j->scene;
	for (int64_t i = begin; i < end; ++i)
	{
		ray_t r;
		r.o.x = j->origins[3 * i]; r.o.y = j->origins[3 * i + 1]; r.o.z = j->origins[3 * i + 2];
		r.d.x = j->directions[3 * i]; r.d.y = j->directions[3 * i + 1]; r.d.z = j->directions[3 * i + 2];
		r.near_ = j->near_far[2 * i]; r.far_ = j->near_far[2 * i + 1];
		float m = 1.0f;
		if (scene->instance_count == 0 || scene->instance_node_count == 0) m = 0.0f; /* ColorF(0.0f), :401 */
		else
		{
			const rzb_node* root = scene->instance_nodes;
			if (box_hit(root->bb_min, root->bb_max, &r, j->minmax) && world_any(scene, 0, &r, j->minmax)) m = 0.0f;
		}
		j->masks_out[4 * i] = j->masks_out[4 * i + 1] = j->masks_out[4 * i + 2] = j->masks_out[4 * i + 3] = m;
	}
}

void rzo_trace_any(const rzb_scene* scene, const float* origins, const float* directions, const float* near_far,
	uint32_t n, int minmax, float* masks_out)
{
	trace_job j;
	j.scene = scene; j.origins = origins; j.directions = directions; j.near_far = near_far;
	j.order = RZO_ORDER_CPU; j.minmax = minmax; j.hits_out = NULL; j.masks_out = masks_out;
	atomic_init(&j.s_top, 0); atomic_init(&j.s_inst, 0); atomic_init(&j.s_mesh, 0); atomic_init(&j.s_tri, 0);
	parallel_for((int64_t)n, any_range, &j);
}

/* Kernel::generateSimpleRay, cpu_engine_kernel.cpp:180-204 */
void rzo_camera_rays(const rzb_camera* cam, float* origins, float* directions, float* near_far)
{
	const float tana = tanf(cam->fov * 0.5f);
	const float aspect = (float)cam->width / (float)cam->height; /* camera.cpp:52 */
	for (uint32_t y = 0; y < cam->height; ++y)
		for (uint32_t x = 0; x < cam->width; ++x)
		{
			const size_t i = (size_t)y * cam->width + x;
			const float dx = ((((float)x + 0.5f) / (float)cam->width) - 0.5f) * tana;
			const float dy = ((((float)y + 0.5f) / (float)cam->height) - 0.5f) * (-tana / aspect);
			/* CoordSystem::transformForward: x_axis * v.x + y_axis * v.y + z_axis * v.z */
			v3 d = {cam->axis_x[0] * dx + cam->axis_y[0] * dy + cam->axis_z[0] * 1.0f,
				cam->axis_x[1] * dx + cam->axis_y[1] * dy + cam->axis_z[1] * 1.0f,
				cam->axis_x[2] * dx + cam->axis_y[2] * dy + cam->axis_z[2] * 1.0f};
			const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
			directions[3 * i] = d.x / len; directions[3 * i + 1] = d.y / len; directions[3 * i + 2] = d.z / len;
			/* origin = transformForward(0) + position */
			const float z = 0.0f;
			origins[3 * i] = (cam->axis_x[0] * z + cam->axis_y[0] * z + cam->axis_z[0] * z) + cam->position[0];
			origins[3 * i + 1] = (cam->axis_x[1] * z + cam->axis_y[1] * z + cam->axis_z[1] * z) + cam->position[1];
			origins[3 * i + 2] = (cam->axis_x[2] * z + cam->axis_y[2] * z + cam->axis_z[2] * z) + cam->position[2];
			near_far[2 * i] = cam->near_far[0];
			near_far[2 * i + 1] = cam->near_far[1];
		}
}

/* cpu_engine_renderer.cpp:224-235; aperture_area = aperture^2 * pi (:196-197) */
void rzo_tonemap(const float* accum, uint32_t n, float aperture, float exposure_time, uint8_t* rgba8)
{
	const float aperture_area = aperture * aperture * 3.14159265358979323846f;
	for (uint32_t i = 0; i < n; ++i)
	{
		const float a = accum[4 * i + 3] == 0.0f ? 1.0f : accum[4 * i + 3];
		for (int k = 0; k < 3; ++k)
		{
			float c = accum[4 * i + k] / a;
			c *= aperture_area;
			c *= exposure_time;
			c *= 1.0e5f;
			c = c / (c + 1.0f);
			rgba8[4 * i + k] = (uint8_t)(c * 255.0f);
		}
		rgba8[4 * i + 3] = 255;
	}
}

int rzo_threads(void) { return pf_thread_count(); }
