// Device-side data layout and the parity-critical primitives (slab, triangle, instance transform) of the B200 render path;
// the traversal itself is in rzb_traverse.cuh.
//
// What is reproduced, bit for bit, from the reference (so that closest-hit IDs match):
//   slab test        BoundingBox::rayIntersection   /root/reference/RayZath/cuda_render_parts.cuh:1178-1191
//   triangle test    Triangle::closestIntersection  /root/reference/RayZath/cuda_render_parts.cuh:1023-1054
//   instance entry   Instance::closestIntersection  /root/reference/RayZath/cuda_instance.cuh:186-214
//   visiting order   Mesh::closestIntersection / ObjectContainerWithBVH::closestIntersection
//                    /root/reference/RayZath/cuda_instance.cuh:35-91, cuda_bvh.cuh:114-171 (near child first
//                    by ray sign on the split axis; the far child's box is tested against the range
//                    AFTER the near subtree has been searched)
// What is NOT taken from the reference: the control structure. The reference walks with a per-thread
// 32-entry local-memory array and re-reads 48-byte nodes one at a time; here sibling pairs are one
// 64-byte aligned fetch (4 x LDG.128), triangles are a 48-byte hot record (3 x LDG.128) with shading
// data split off, the far child is deferred on a short stack in shared memory together with its slab
// entry distance (so the late range test is a compare, not a second node fetch), and both BVH levels
// share one flat loop. All parity-critical arithmetic uses __f*_rn intrinsics: no FMA contraction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rzb200.h"

namespace rzb
{
	constexpr uint32_t kNoIndex = 0xFFFFFFFFu;
	constexpr float kFltMax = 3.402823466e+38f;

	// ---- parity-critical scalar ops: round-to-nearest, never contracted ----
	__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
	__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
	__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
	__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

	struct V3
	{
		float x, y, z;
	};
	__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
	// vec3f::dotProduct / crossProduct (cuda_render_parts.cuh:46-57): left-to-right, no FMA
	__device__ __forceinline__ float dot_rn(const V3& a, const V3& b)
	{
		return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
	}
	__device__ __forceinline__ V3 cross_rn(const V3& a, const V3& b)
	{
		return v3(
			fsub(fmul(a.y, b.z), fmul(a.z, b.y)),
			fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
			fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
	}
	__device__ __forceinline__ V3 sub_rn(const V3& a, const V3& b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }

	// ---- device scene ----
	// Instance record for traversal, 96 B = 6 x float4.
	struct __align__(16) DInstance
	{
		float px, py, pz, sx;
		float sy, sz, xx, xy;
		float xz, yx, yy, yz;
		float zx, zy, zz, bminx;
		float bminy, bminz, bmaxx, bmaxy;
		float bmaxz;
		uint32_t mesh_root; // global node index of the mesh root, kNoIndex = no mesh / empty mesh
		uint32_t mat_offset;
		uint32_t mat_count;
	};
	static_assert(sizeof(DInstance) == 96, "DInstance layout");

	struct DMap
	{
		const void* pixels;
		uint32_t width, height;
		uint32_t format, filter, address;
		float scale_x, scale_y;
		float rot_sin, rot_cos;
		float trans_x, trans_y;
		uint32_t _pad;
	};

	struct DScene
	{
		const float4* nodes;    // 2 x float4 per node: {min.xyz, max.x}, {max.y, max.z, begin, type_count}; global indices
		const float4* nodes4;   // wide (4-ary) mesh nodes, 8 x float4 each (RZB_SCENE_WIDE_TREES), else NULL
		const uint32_t* inst_root4; // per instance: reference of its mesh's wide root (rzb_traverse.cuh: wide_decode)
		const float4* tri_hot;  // 3 x float4 per triangle: {v1.xyz, e1.x}, {e1.yz, e2.xy}, {e2.z, slot, -, -}
		const float4* tri_cold; // 5 x float4 per triangle: n1, n2, n3, face normal, uvs
		const DInstance* instances;
		const uint32_t* inst_host_index;
		const uint32_t* tri_host_index;
		const uint32_t* inst_materials;
		const rzb_material* materials; // [material_count] + world material at index world_material
		const DMap* maps;
		const rzb_direct_light* direct_lights;
		const rzb_spot_light* spot_lights;
		uint32_t top_root;       // global node index of the instance tree root
		uint32_t instance_count;
		uint32_t material_count;
		uint32_t world_material; // index of the world material in `materials`
		uint32_t default_material;
		uint32_t direct_light_count, spot_light_count;
		uint32_t flags;
	};

	struct Hit
	{
		float t;       // ray.near_far.y after traversal
		float near_;   // ray.near_far.x after traversal (changes when an instance registers a hit)
		float b1, b2;
		uint32_t tri;  // global triangle index (BVH order), kNoIndex on miss
		uint32_t inst; // index into the BVH-ordered instance array
		bool external;
	};

	struct TraceCounters
	{
		uint32_t top_nodes, instances, mesh_nodes, triangles;
	};

	// ---- short stack: first kSmemStack entries per thread in shared memory (interleaved by thread so a
	// warp's pushes hit 32 consecutive 8-byte words), the rest in local memory. Depth bound: two trees of
	// depth <= 33 (max_depth 31 in both builders) plus one instance-range entry. rzb_set_scene REJECTS scenes whose
	// trees could need more than kSmemStack + kLocalStack entries (depth of the instance tree + 1 + depth of the deepest
	// mesh tree, computed on the host; children must follow their parent, so trees are acyclic), so push never
	// overflows for an accepted scene; the index clamps below only keep a corrupted tree from reading out of bounds.
	// Shared memory taken here is L1 taken from the node / triangle fetches (one 256 KB array per SM): measured on B200
	// at 1080p, entries in shared memory 20 / 12 / 8 / 4 / 2 -> materials scene 1058 / 1100 / 1108 / 1103 / 1088 Mrays/s,
	// 1M triangles 1140 / 1159 / 1161 / 1153 / 1135 (at 8 blocks per SM the 20-entry stack left the shadow kernel
	// almost no L1).
#ifndef RZB_SMEM_STACK
#define RZB_SMEM_STACK 8
#endif
	constexpr int kSmemStack = RZB_SMEM_STACK;
	constexpr int kLocalStack = 64;
	constexpr int kTraceBlock = 128;

	enum : uint32_t
	{
		kEntryMeshNode = 0u << 30,
		kEntryTopNode = 1u << 30,
		kEntryInstRange = 2u << 30,
		kEntryKindMask = 3u << 30,
		kEntryIndexMask = ~(3u << 30)
	};

	struct Stack
	{
		// Shared-memory part: &smem_stack[threadIdx.x], stride kTraceBlock entries. On the device it is kept as a 32-bit
		// shared-window address and accessed with st.shared / ld.shared: as a generic pointer inside this struct the compiler
		// emitted generic ST.E.64 / LD.E.64 with 64-bit address arithmetic (two registers for the base) for every push and pop.
#ifdef __CUDA_ARCH__
		uint32_t smem;
#else
		uint2* smem;
#endif
		uint2 local[kLocalStack];
		int sp;
		static __device__ __forceinline__ int clamp_local(int i) { return i < kLocalStack ? i : kLocalStack - 1; }
		__device__ __forceinline__ void set_smem(uint2* base_for_thread)
		{
#ifdef __CUDA_ARCH__
			smem = static_cast<uint32_t>(__cvta_generic_to_shared(base_for_thread));
#else
			smem = base_for_thread;
#endif
		}
		__device__ __forceinline__ void smem_store(const int i, const uint32_t a, const uint32_t b)
		{
#ifdef __CUDA_ARCH__
			asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem + uint32_t(i) * uint32_t(kTraceBlock * 8)), "r"(a), "r"(b) : "memory");
#else
			smem[i * kTraceBlock] = make_uint2(a, b);
#endif
		}
		__device__ __forceinline__ uint2 smem_load(const int i) const
		{
#ifdef __CUDA_ARCH__
			uint2 r;
			asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(smem + uint32_t(i) * uint32_t(kTraceBlock * 8)) : "memory");
			return r;
#else
			return smem[i * kTraceBlock];
#endif
		}
		__device__ __forceinline__ void push(uint32_t a, uint32_t b)
		{
			if (sp < kSmemStack) smem_store(sp, a, b);
			else local[clamp_local(sp - kSmemStack)] = make_uint2(a, b);
			++sp;
		}
		__device__ __forceinline__ uint2 pop()
		{
			--sp;
			if (sp < kSmemStack) return smem_load(sp);
			return local[clamp_local(sp - kSmemStack)];
		}
		__device__ __forceinline__ uint2 peek() const
		{
			if (sp - 1 < kSmemStack) return smem_load(sp - 1);
			return local[clamp_local(sp - 1 - kSmemStack)];
		}
	};

	// BoundingBox::rayIntersection: six IEEE divides, fminf/fmaxf, then the three range tests.
	// Returns the far-independent part; tmin is handed back for the (possibly deferred) `tmin > far` test.
	__device__ __forceinline__ bool slab_rn(const float4 n0, const float4 n1, const V3& o, const V3& d,
		const float near_, float& tmin)
	{
		const float t1 = fdiv(fsub(n0.x, o.x), d.x);
		const float t2 = fdiv(fsub(n0.w, o.x), d.x);
		const float t3 = fdiv(fsub(n0.y, o.y), d.y);
		const float t4 = fdiv(fsub(n1.x, o.y), d.y);
		const float t5 = fdiv(fsub(n0.z, o.z), d.z);
		const float t6 = fdiv(fsub(n1.y, o.z), d.z);
		tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
		const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
		return !(tmax < near_ || tmin > tmax);
	}
	// the reference's full predicate is  !(tmax < near || tmin > tmax || tmin > far)
	__device__ __forceinline__ bool range_ok(const float tmin, const float far_) { return !(tmin > far_); }

	// Triangle::closestIntersection on the hot record. Returns true when the hit was accepted.
	__device__ __forceinline__ bool triangle_closest(const float4* __restrict__ tri_hot, const uint32_t tri,
		const V3& o, const V3& d, const float near_, float& far_, float& b1_out, float& b2_out, bool& external)
	{
		const float4 q0 = __ldg(tri_hot + 3 * size_t(tri));
		const float4 q1 = __ldg(tri_hot + 3 * size_t(tri) + 1);
		const float4 q2 = __ldg(tri_hot + 3 * size_t(tri) + 2);
		const V3 v1 = v3(q0.x, q0.y, q0.z);
		const V3 e1 = v3(q0.w, q1.x, q1.y);
		const V3 e2 = v3(q1.z, q1.w, q2.x);
		const V3 pvec = cross_rn(d, e2);
		float det = dot_rn(e1, pvec);
		if (det > -1.0e-7f && det < 1.0e-7f) det = fadd(det, 1.0e-7f);
		const float inv_det = fdiv(1.0f, det);
		const V3 tvec = sub_rn(o, v1);
		const float b1 = fmul(dot_rn(tvec, pvec), inv_det);
		if (b1 < 0.0f || b1 > 1.0f) return false;
		const V3 qvec = cross_rn(tvec, e1);
		const float b2 = fmul(dot_rn(d, qvec), inv_det);
		if (b2 < 0.0f || fadd(b1, b2) > 1.0f) return false;
		const float t = fmul(dot_rn(e2, qvec), inv_det);
		if (t <= near_ || t >= far_) return false;
		far_ = t;
		b1_out = b1;
		b2_out = b2;
		external = det > 0.0f;
		return true;
	}

	// Transformation::transformG2L + length factor (cuda_instance.cuh:193-199, cuda_render_parts.cuh:1150-1158).
	// Normalisation divides by the magnitude (the host Math library's Normalize as restated in oracle/shim/vec3.h;
	// the reference's CUDA side uses rnorm3df — see DESIGN.md "normalisation").
	__device__ __forceinline__ void ray_to_local(const DInstance& in, const V3& wo, const V3& wd,
		V3& lo, V3& ld, float& len)
	{
		const V3 p = sub_rn(wo, v3(in.px, in.py, in.pz));
		lo = v3(
			fdiv(fadd(fadd(fmul(in.xx, p.x), fmul(in.xy, p.y)), fmul(in.xz, p.z)), in.sx),
			fdiv(fadd(fadd(fmul(in.yx, p.x), fmul(in.yy, p.y)), fmul(in.yz, p.z)), in.sy),
			fdiv(fadd(fadd(fmul(in.zx, p.x), fmul(in.zy, p.y)), fmul(in.zz, p.z)), in.sz));
		ld = v3(
			fdiv(fadd(fadd(fmul(in.xx, wd.x), fmul(in.xy, wd.y)), fmul(in.xz, wd.z)), in.sx),
			fdiv(fadd(fadd(fmul(in.yx, wd.x), fmul(in.yy, wd.y)), fmul(in.yz, wd.z)), in.sy),
			fdiv(fadd(fadd(fmul(in.zx, wd.x), fmul(in.zy, wd.y)), fmul(in.zz, wd.z)), in.sz));
		len = __fsqrt_rn(fadd(fadd(fmul(ld.x, ld.x), fmul(ld.y, ld.y)), fmul(ld.z, ld.z)));
		ld = v3(fdiv(ld.x, len), fdiv(ld.y, len), fdiv(ld.z, len));
	}

	__device__ __forceinline__ uint32_t sign_bits(const V3& d)
	{
		// bit 2 = x, bit 1 = y, bit 0 = z (cuda_instance.cuh:49-52); split types X=2, Y=1, Z=0, Size=3
		return (uint32_t(d.x < 0.0f) << 2) | (uint32_t(d.y < 0.0f) << 1) | uint32_t(d.z < 0.0f);
	}

	__device__ __forceinline__ DInstance load_instance(const DInstance* __restrict__ instances, uint32_t i)
	{
		const float4* p = reinterpret_cast<const float4*>(instances + i);
		DInstance r;
		float4* q = reinterpret_cast<float4*>(&r);
#pragma unroll
		for (int k = 0; k < 6; ++k) q[k] = __ldg(p + k);
		return r;
	}
}
