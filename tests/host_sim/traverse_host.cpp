// Host build of the device traversal (rayzath_b200/csrc/rzb_traverse.cuh) for logic tests on CPU: the CUDA
// intrinsics are mapped to plain fp32 operations (compile with -ffp-contract=off), one "lane" at a time.
// Test-only: checks the control flow (phases, deferred stack, instance transitions, work counters) against the
// oracle without a GPU. Exposes trav_host_closest() with the C-ABI scene.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs(int(v)); }
struct { unsigned x = 0; } threadIdx;
#define RZB_HOST_SIM 1
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#include "../../rayzath_b200/csrc/rzb_traverse.cuh"
#include "../../rayzath_b200/csrc/rzb_traverse_mr.cuh"
#include "../../rayzath_b200/csrc/rzb_wide.hpp"

using namespace rzb;

// the real shadow_attenuation lives in rzb_shade.cuh; the host simulation only runs CPU-semantics shadows
namespace rzb
{
	float4 shadow_attenuation(const DScene&, const uint32_t, const float, const float, const uint32_t, const uint32_t)
	{
		return make_float4(0, 0, 0, 0);
	}
}

static int trav_host_run_impl(const rzb_scene* s, const float* origins, const float* dirs, const float* near_far, uint32_t n,
	int any, rzb_hit* hits_out, float* masks_out, uint64_t* counters4, int mr);

extern "C" int trav_host_run(const rzb_scene* s, const float* origins, const float* dirs, const float* near_far, uint32_t n,
	int any, rzb_hit* hits_out, float* masks_out, uint64_t* counters4)
{
	return trav_host_run_impl(s, origins, dirs, near_far, n, any, hits_out, masks_out, counters4, 0);
}
// the own-tree kernels' walk on 4-ary trees (RZB_SCENE_WIDE_TREES): conservative boxes, wide collapse of the given trees
extern "C" int trav_host_run_wide(const rzb_scene* s, const float* origins, const float* dirs, const float* near_far, uint32_t n,
	rzb_hit* hits_out, uint64_t* counters4)
{
	return trav_host_run_impl(s, origins, dirs, near_far, n, 0, hits_out, nullptr, counters4, 2);
}
// the multi-ray-per-lane walk (rzb_traverse_mr.cuh): ONE lane with kMrRays rays in flight, phases chosen by mr_vote
extern "C" int trav_host_run_mr(const rzb_scene* s, const float* origins, const float* dirs, const float* near_far, uint32_t n,
	rzb_hit* hits_out, uint64_t* counters4)
{
	return trav_host_run_impl(s, origins, dirs, near_far, n, 0, hits_out, nullptr, counters4, 1);
}

static int trav_host_run_impl(const rzb_scene* s, const float* origins, const float* dirs, const float* near_far, uint32_t n,
	int any, rzb_hit* hits_out, float* masks_out, uint64_t* counters4, int mr)
{
	// build the device layout exactly like rzb_set_scene does (single mesh table walk, host side)
	std::vector<uint32_t> mesh_base(s->mesh_count, kNoIndex);
	size_t cursor = 1;
	for (uint32_t m = 0; m < s->mesh_count; ++m)
	{
		if (s->meshes[m].node_count == 0) continue;
		if ((cursor & 1u) == 0) ++cursor;
		mesh_base[m] = uint32_t(cursor);
		cursor += s->meshes[m].node_count;
	}
	if ((cursor & 1u) == 0) ++cursor;
	const uint32_t top_base = uint32_t(cursor);
	cursor += s->instance_node_count;
	std::vector<rzb_node> nodes(cursor + 1);
	std::memset(nodes.data(), 0, nodes.size() * sizeof(rzb_node));
	for (uint32_t m = 0; m < s->mesh_count; ++m)
		for (uint32_t i = 0; i < s->meshes[m].node_count; ++i)
		{
			rzb_node nd = s->mesh_nodes[s->meshes[m].node_offset + i];
			nd.begin += (nd.type_count & 0x3FFFFFFFu) ? s->meshes[m].tri_offset : mesh_base[m];
			nodes[mesh_base[m] + i] = nd;
		}
	for (uint32_t i = 0; i < s->instance_node_count; ++i)
	{
		rzb_node nd = s->instance_nodes[i];
		if ((nd.type_count & 0x3FFFFFFFu) == 0) nd.begin += top_base;
		nodes[top_base + i] = nd;
	}
	std::vector<float4> hot(size_t(s->triangle_count) * 3 + 1);
	for (uint32_t i = 0; i < s->triangle_count; ++i)
	{
		const rzb_triangle& t = s->triangles[i];
		const float e1x = t.v[1][0] - t.v[0][0], e1y = t.v[1][1] - t.v[0][1], e1z = t.v[1][2] - t.v[0][2];
		const float e2x = t.v[2][0] - t.v[0][0], e2y = t.v[2][1] - t.v[0][1], e2z = t.v[2][2] - t.v[0][2];
		hot[3 * size_t(i)] = make_float4(t.v[0][0], t.v[0][1], t.v[0][2], e1x);
		hot[3 * size_t(i) + 1] = make_float4(e1y, e1z, e2x, e2y);
		hot[3 * size_t(i) + 2] = make_float4(e2z, __uint_as_float(t.material_slot & 0x3Fu), 0.0f, 0.0f);
	}
	std::vector<DInstance> insts(s->instance_count + 1);
	for (uint32_t i = 0; i < s->instance_count; ++i)
	{
		const rzb_instance& h = s->instances[i];
		DInstance d{};
		d.px = h.position[0]; d.py = h.position[1]; d.pz = h.position[2];
		d.sx = h.scale[0]; d.sy = h.scale[1]; d.sz = h.scale[2];
		d.xx = h.axis_x[0]; d.xy = h.axis_x[1]; d.xz = h.axis_x[2];
		d.yx = h.axis_y[0]; d.yy = h.axis_y[1]; d.yz = h.axis_y[2];
		d.zx = h.axis_z[0]; d.zy = h.axis_z[1]; d.zz = h.axis_z[2];
		d.bminx = h.bb_min[0]; d.bminy = h.bb_min[1]; d.bminz = h.bb_min[2];
		d.bmaxx = h.bb_max[0]; d.bmaxy = h.bb_max[1]; d.bmaxz = h.bb_max[2];
		d.mesh_root = h.mesh == RZB_NO_INDEX ? kNoIndex : mesh_base[h.mesh];
		d.mat_offset = h.material_offset; d.mat_count = h.material_count;
		insts[i] = d;
	}
	DScene sc{};
	sc.nodes = reinterpret_cast<const float4*>(nodes.data());
	sc.tri_hot = hot.data();
	sc.instances = insts.data();
	sc.top_root = top_base;
	sc.instance_count = s->instance_count;
	sc.flags = RZB_FLAG_CPU_SEMANTICS;

	std::vector<WideNode> wide;
	std::vector<uint32_t> inst_root4(s->instance_count + 1, kWideEmpty);
	if (mr == 2)
	{
		std::vector<uint32_t> root4(s->mesh_count, kWideEmpty);
		uint32_t depth = 0;
		bool ok = true;
		for (uint32_t m = 0; m < s->mesh_count; ++m)
			if (mesh_base[m] != kNoIndex)
				root4[m] = collapseWide(s->mesh_nodes + s->meshes[m].node_offset, 0u, s->meshes[m].tri_offset, wide, 0u, depth, ok);
		if (!ok) return 1;
		for (uint32_t i = 0; i < s->instance_count; ++i)
			if (s->instances[i].mesh != RZB_NO_INDEX) inst_root4[i] = root4[s->instances[i].mesh];
		wide.emplace_back();
		sc.nodes4 = reinterpret_cast<const float4*>(wide.data());
		sc.inst_root4 = inst_root4.data();
	}
	if (mr == 1)
	{
		constexpr int K = kMrMaxRays;
		std::vector<float4> smem(size_t(kMrFields) * K * kMrBlock);
		MrHot<K> hot{smem.data()};
		static MrCold<K> cold;
		MrLane lane;
		lane.init(K);
		TraceCounters cnt{0u, 0u, 0u, 0u};
		uint32_t next_ray = 0;
		auto write = [&](int k) {
			RayResult r;
			mr_result<K>(hot, cold, k, r);
			const uint32_t i = cold.handle[k];
			rzb_hit h{};
			h.instance = RZB_NO_INDEX; h.triangle = RZB_NO_INDEX;
			h.t = r.t;
			if (r.inst != kNoIndex)
			{
				h.instance = s->instances[r.inst].host_index;
				h.triangle = s->tri_host_index ? s->tri_host_index[r.tri] : r.tri;
				h.b1 = r.b1; h.b2 = r.b2; h.external = r.external ? 1u : 0u;
			}
			hits_out[i] = h;
		};
		for (;;)
		{
			const bool work_left = next_ray < n;
			const uint32_t phase = mr_vote(lane, work_left);
			if (phase == kMrDead) break;
			if (phase == kMrNode) mr_node<K, false, true, 2>(sc, hot, cold, lane, lane.pick(kMrNode), cnt);
			else if (phase == kMrLeaf) mr_leaf<K, false, true>(sc, hot, cold, lane, lane.pick(kMrLeaf), cnt);
			else if (phase == kMrHeavy) mr_heavy<K, false, true>(sc, hot, cold, lane, lane.pick(kMrHeavy), cnt);
			else
			{
				int k = lane.pick(kMrDone);
				if (k >= 0) { write(k); lane.move(k, kMrDone, kMrEmpty); }
				else if (work_left) k = lane.pick(kMrEmpty);
				if (k >= 0 && work_left)
				{
					const uint32_t i = next_ray++;
					cold.handle[k] = i; cold.user[k] = 0u;
					mr_begin<K, false, true>(sc, hot, cold, lane, k, v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
						v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), near_far[2 * i], near_far[2 * i + 1], cnt);
				}
			}
			lane.rr = (lane.rr + 1u) & 3u;
		}
		if (counters4)
		{
			counters4[0] = cnt.top_nodes; counters4[1] = cnt.instances; counters4[2] = cnt.mesh_nodes; counters4[3] = cnt.triangles;
		}
		return 0;
	}
	std::vector<uint2> smem(size_t(kSmemStack) * kTraceBlock);
	Stack st;
	st.set_smem(smem.data());
	st.sp = 0;
	TraceCounters cnt{0u, 0u, 0u, 0u};
	uint64_t total[4] = {0, 0, 0, 0};
	for (uint32_t i = 0; i < n; ++i)
	{
		const V3 o = v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
		const V3 d = v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
		cnt = TraceCounters{0u, 0u, 0u, 0u};
		ParkedRay park;
		RayResult r;
		if (any)
		{
			trace_ray<true, true>(sc, true, o, d, near_far[2 * i], near_far[2 * i + 1], st, park, cnt, r);
			masks_out[4 * i] = r.mask.x; masks_out[4 * i + 1] = r.mask.y; masks_out[4 * i + 2] = r.mask.z; masks_out[4 * i + 3] = r.mask.w;
		}
		else
		{
			if (mr == 2) trace_ray<false, true, 0, true, true>(sc, true, o, d, near_far[2 * i], near_far[2 * i + 1], st, park, cnt, r);
			else trace_ray<false, true>(sc, true, o, d, near_far[2 * i], near_far[2 * i + 1], st, park, cnt, r);
			rzb_hit h{};
			h.instance = RZB_NO_INDEX; h.triangle = RZB_NO_INDEX;
			h.t = r.t;
			if (r.inst != kNoIndex)
			{
				h.instance = s->instances[r.inst].host_index;
				h.triangle = s->tri_host_index ? s->tri_host_index[r.tri] : r.tri;
				h.b1 = r.b1; h.b2 = r.b2; h.external = r.external ? 1u : 0u;
			}
			hits_out[i] = h;
		}
		total[0] += cnt.top_nodes; total[1] += cnt.instances; total[2] += cnt.mesh_nodes; total[3] += cnt.triangles;
	}
	if (counters4) std::memcpy(counters4, total, sizeof(total));
	return 0;
}
