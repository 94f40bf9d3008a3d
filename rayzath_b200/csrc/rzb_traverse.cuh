// Two-level BVH traversal of the B200 render path: closest hit and any hit (shadow) on the reference's own trees.
//
// Decisions are bit-identical to the reference's arithmetic (so closest-hit ids match, tests/test_gpu_parity.py):
//   slab test        BoundingBox::rayIntersection   /root/reference/RayZath/cuda_render_parts.cuh:1178-1191
//   triangle test    Triangle::closestIntersection  /root/reference/RayZath/cuda_render_parts.cuh:1023-1054
//   instance entry   Instance::closestIntersection  /root/reference/RayZath/cuda_instance.cuh:186-214
//   visiting order   Mesh::closestIntersection / ObjectContainerWithBVH::closestIntersection
//                    /root/reference/RayZath/cuda_instance.cuh:35-91, cuda_bvh.cuh:114-171 (near child first by ray
//                    sign on the split axis; the far child sees the range as it is AFTER the near subtree)
// The control structure is not the reference's:
//   * sibling pairs are one 64-byte aligned fetch (4 x LDG.128); triangles are a 48-byte hot record (3 x LDG.128);
//   * the six IEEE divisions of a slab test are replaced by multiplications with the reciprocal direction and an
//     error margin: the product differs from the correctly rounded quotient by < 2 ulp, so whenever the three
//     comparisons of the predicate are decided by more than that margin the decision equals the reference's;
//     otherwise (about one test in 10^5) the exact divisions are evaluated. ncu on the first version of this
//     kernel showed 38 % of all issued instructions inside division sequences (profiles/r01_*);
//   * the far child is deferred on a short stack (shared memory, interleaved by lane) together with its entry
//     distance, so the reference's late range test is a compare at pop time instead of a second box test;
//   * traversal is phase-structured (inner nodes / leaf triangles / instance transitions) so that the lanes of a
//     warp run the same phase together, and lanes whose ray has finished pull a new ray immediately
//     (per-warp ballot + one atomic) instead of idling until the slowest lane of the warp is done.
#pragma once

#include "rzb_device.cuh"

namespace rzb
{
	constexpr float kSlabMargin = 4.0e-7f; // > 2 ulp relative (2^-22 = 2.4e-7)

	struct SlabResult
	{
		bool box;   // !(tmax < near || tmin > tmax)
		bool range; // !(tmin > far)
		float tmin; // entry distance (approximate unless the exact path ran)
	};

	// exact predicate (the reference's arithmetic)
#ifndef RZB_SLAB_EXACT_INLINE
#define RZB_SLAB_EXACT_ATTR __noinline__
#else
#define RZB_SLAB_EXACT_ATTR __forceinline__
#endif
	__device__ RZB_SLAB_EXACT_ATTR SlabResult slab_exact(const float4 n0, const float4 n1, const V3 o, const V3 d,
		const float near_, const float far_)
	{
		SlabResult r;
		r.box = slab_rn(n0, n1, o, d, near_, r.tmin);
		r.range = range_ok(r.tmin, far_);
		return r;
	}

	// fast predicate with fallback; `rcp` = 1/d per component (IEEE), `exact_only` forces the fallback (denormal d)
	__device__ __forceinline__ SlabResult slab_fast(const float4 n0, const float4 n1, const V3& o, const V3& d, const V3& rcp,
		const float near_, const float far_, const bool exact_only)
	{
		const float t1 = fmul(fsub(n0.x, o.x), rcp.x);
		const float t2 = fmul(fsub(n0.w, o.x), rcp.x);
		const float t3 = fmul(fsub(n0.y, o.y), rcp.y);
		const float t4 = fmul(fsub(n1.x, o.y), rcp.y);
		const float t5 = fmul(fsub(n0.z, o.z), rcp.z);
		const float t6 = fmul(fsub(n1.y, o.z), rcp.z);
		const float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
		const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
		// margins relative to the approximate values themselves (near/far are exact operands); infinities
		// (a direction component == 0) are exact in both formulations and never "close"
		const float m_min = fmaf(kSlabMargin, fminf(fabsf(tmin), 1.0e30f), 1.0e-37f);
		const float m_max = fmaf(kSlabMargin, fminf(fabsf(tmax), 1.0e30f), 1.0e-37f);
		const bool ambiguous = exact_only ||
			fabsf(tmax - near_) <= m_max || fabsf(tmin - tmax) <= fmaxf(m_min, m_max) || fabsf(tmin - far_) <= m_min;
		if (ambiguous) return slab_exact(n0, n1, o, d, near_, far_);
		SlabResult r;
		r.box = !(tmax < near_ || tmin > tmax);
		r.range = !(tmin > far_);
		r.tmin = tmin;
		return r;
	}

	__device__ __forceinline__ V3 reciprocal_rn(const V3& d) { return v3(fdiv(1.0f, d.x), fdiv(1.0f, d.y), fdiv(1.0f, d.z)); }
	// a direction component so small that its reciprocal overflows while quotients may not: use exact divisions
	__device__ __forceinline__ bool needs_exact(const V3& d)
	{
		const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
		return (ax != 0.0f && ax < 1.0e-30f) || (ay != 0.0f && ay < 1.0e-30f) || (az != 0.0f && az < 1.0e-30f);
	}

	enum : uint32_t { kTravNode = 0u, kTravPop = 1u, kTravDone = 2u };

	// Per-lane traversal state. MODE_ANY = shadow query (fixed child order, mask product, early out).
	template <bool ANY>
	struct Traversal
	{
		V3 wo, wd;            // world ray
		float wnear, wfar;    // world range (far shrinks with every registered hit)
		V3 o, d, rcp;         // current level
		float near_, far_, len;
		uint32_t sbits;
		uint32_t cur_begin, cur_tc;
		uint32_t state;       // kTrav*
		uint32_t cur_inst;
		bool in_mesh, mesh_hit, exact_only;
		// closest-hit result
		uint32_t hit_inst, hit_tri, ltri;
		float b1, b2, lb1, lb2;
		bool ext, lext;
		// any-hit result
		float4 mask;
		uint32_t mat_offset, mat_count;

		__device__ __forceinline__ void to_world_level()
		{
			o = wo; d = wd;
			rcp = reciprocal_rn(wd);
			exact_only = needs_exact(wd);
			sbits = ANY ? 0u : sign_bits(wd);
			near_ = wnear; far_ = wfar;
			len = 1.0f;
			in_mesh = false;
		}

		// start a ray: root test of the instance tree
		template <bool STATS>
		__device__ __forceinline__ void begin(const DScene& sc, const V3 origin, const V3 dir, const float n, const float f,
			Stack& st, TraceCounters& cnt)
		{
			wo = origin; wd = dir; wnear = n; wfar = f;
			hit_inst = kNoIndex; hit_tri = kNoIndex; b1 = 0.0f; b2 = 0.0f; ext = true;
			ltri = kNoIndex; lb1 = lb2 = 0.0f; lext = true;
			mask = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
			mat_offset = mat_count = 0u;
			mesh_hit = false; cur_inst = kNoIndex;
			st.sp = 0;
			to_world_level();
			state = kTravDone;
			if (sc.instance_count == 0u)
			{
				// no instances: the CPU engine's shadow query answers "occluded" (cpu_engine_kernel.cpp:401), the CUDA one "free"
				if (ANY && (sc.flags & RZB_FLAG_CPU_SEMANTICS)) mask = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				return;
			}
			const float4 n0 = __ldg(sc.nodes + 2 * size_t(sc.top_root));
			const float4 n1 = __ldg(sc.nodes + 2 * size_t(sc.top_root) + 1);
			if (STATS) cnt.top_nodes++;
			const SlabResult r = slab_fast(n0, n1, o, d, rcp, near_, far_, exact_only);
			if (!(r.box && r.range)) return;
			cur_begin = __float_as_uint(n1.z);
			cur_tc = __float_as_uint(n1.w);
			state = kTravNode;
		}

		// Phase 1: walk inner nodes (and pop deferred nodes) until a leaf is current, or something other than a node
		// of the current level has to be popped (state = kTravPop with the entry left on the stack), or the stack is empty.
		template <bool STATS>
		__device__ __forceinline__ void inner_phase(const DScene& sc, Stack& st, TraceCounters& cnt)
		{
			const float4* __restrict__ nodes = sc.nodes;
			if (state == kTravDone) return;
			for (;;)
			{
				if (state == kTravPop)
				{
					if (st.sp == 0) return;
					const uint2 e = st.peek();
					const uint32_t kind = e.x & kEntryKindMask;
					if (kind == kEntryInstRange || (in_mesh && kind != kEntryMeshNode)) return; // transition phase
					st.sp--;
					const uint32_t idx = e.x & kEntryIndexMask;
					if (!ANY)
					{
						// late range test of a deferred node (the reference tests the far child after the near subtree)
						const float tmin = __uint_as_float(e.y);
						const float m = fmaf(kSlabMargin, fminf(fabsf(tmin), 1.0e30f), 1.0e-37f);
						if (tmin > far_ + m) continue;
						if (!(tmin < far_ - m))
						{
							// too close to call with the approximate entry distance: evaluate the reference's arithmetic
							const float4 x0 = __ldg(nodes + 2 * size_t(idx));
							const float4 x1 = __ldg(nodes + 2 * size_t(idx) + 1);
							const SlabResult r = slab_exact(x0, x1, o, d, near_, far_);
							if (!r.range) continue;
						}
					}
					const float4 n1 = __ldg(nodes + 2 * size_t(idx) + 1);
					cur_begin = __float_as_uint(n1.z);
					cur_tc = __float_as_uint(n1.w);
					state = kTravNode;
				}
				if ((cur_tc & 0x3FFFFFFFu) != 0u) return; // leaf
				// inner node: fetch the sibling pair (64 B, 64-byte aligned)
				const uint32_t flip = ANY ? 0u : ((sbits >> (cur_tc >> 30)) & 1u);
				const uint32_t ia = cur_begin + flip, ib = cur_begin + (flip ^ 1u);
				const float4 a0 = __ldg(nodes + 2 * size_t(ia));
				const float4 a1 = __ldg(nodes + 2 * size_t(ia) + 1);
				const float4 c0 = __ldg(nodes + 2 * size_t(ib));
				const float4 c1 = __ldg(nodes + 2 * size_t(ib) + 1);
				if (STATS) { if (in_mesh) cnt.mesh_nodes += 2; else cnt.top_nodes += 2; }
				const SlabResult ra = slab_fast(a0, a1, o, d, rcp, near_, far_, exact_only);
				const SlabResult rb = slab_fast(c0, c1, o, d, rcp, near_, far_, exact_only);
				const bool hit_a = ra.box && ra.range, hit_b = rb.box && rb.range;
				if (hit_a)
				{
					// B is deferred with its entry distance: it is range-tested again when popped, i.e. after A's subtree
					if (hit_b) st.push((in_mesh ? kEntryMeshNode : kEntryTopNode) | ib, __float_as_uint(rb.tmin));
					cur_begin = __float_as_uint(a1.z);
					cur_tc = __float_as_uint(a1.w);
				}
				else if (hit_b)
				{
					cur_begin = __float_as_uint(c1.z);
					cur_tc = __float_as_uint(c1.w);
				}
				else state = kTravPop;
			}
		}

		// Phase 2: the current node is a leaf.
		template <bool STATS>
		__device__ __forceinline__ void leaf_phase(const DScene& sc, Stack& st, TraceCounters& cnt)
		{
			if (state != kTravNode) return;
			const uint32_t count = cur_tc & 0x3FFFFFFFu;
			if (count == 0u) return;
			state = kTravPop;
			if (!in_mesh)
			{
				st.push(kEntryInstRange | cur_begin, cur_begin + count);
				return;
			}
			const uint32_t end = cur_begin + count;
			for (uint32_t i = cur_begin; i < end; ++i)
			{
				if (STATS) cnt.triangles++;
				if (!ANY)
				{
					if (triangle_closest(sc.tri_hot, i, o, d, near_, far_, lb1, lb2, lext))
					{
						ltri = i;
						mesh_hit = true;
					}
				}
				else
				{
					float tf = far_, tb1, tb2;
					bool text;
					if (!triangle_closest(sc.tri_hot, i, o, d, near_, tf, tb1, tb2, text)) continue;
					if (sc.flags & RZB_FLAG_CPU_SEMANTICS) mask = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
					else shadow_attenuate(sc, i, tb1, tb2);
					if (mask.w < 1.0e-4f)
					{
						state = kTravDone;
						return;
					}
				}
			}
		}

		// every intersected triangle multiplies the mask by its material's opacity colour (cuda_instance.cuh:105-112)
		__device__ __forceinline__ void shadow_attenuate(const DScene& sc, const uint32_t i, const float tb1, const float tb2);

		// Phase 3: leave a finished mesh, enter the next instance of a range, or finish the ray.
		template <bool STATS>
		__device__ __forceinline__ void transition_phase(const DScene& sc, Stack& st, TraceCounters& cnt)
		{
			if (state != kTravPop) return;
			uint2 e = make_uint2(kEntryTopNode, 0u);
			const bool have = st.sp != 0;
			if (have) e = st.peek();
			const uint32_t kind = e.x & kEntryKindMask;
			if (in_mesh && (!have || kind != kEntryMeshNode))
			{
				// the current mesh is exhausted: leave the instance (cuda_instance.cuh:203-213)
				if (!ANY && mesh_hit)
				{
					hit_inst = cur_inst; hit_tri = ltri; b1 = lb1; b2 = lb2; ext = lext;
					wnear = fdiv(near_, len);
					wfar = fdiv(far_, len);
				}
				to_world_level();
			}
			if (!have)
			{
				state = kTravDone;
				return;
			}
			if (kind != kEntryInstRange) return; // a deferred top-level node: the inner phase pops it
			st.sp--;
			const uint32_t idx = e.x & kEntryIndexMask, end = e.y;
			if (idx + 1u < end) st.push(kEntryInstRange | (idx + 1u), end);
			// Instance::closestIntersection / anyIntersection (cuda_instance.cuh:186-229)
			if (STATS) cnt.instances++;
			const DInstance in = load_instance(sc.instances, idx);
			const float4 n0 = make_float4(in.bminx, in.bminy, in.bminz, in.bmaxx);
			const float4 n1 = make_float4(in.bmaxy, in.bmaxz, 0.0f, 0.0f);
			const SlabResult rb = slab_fast(n0, n1, o, d, rcp, near_, far_, exact_only);
			if (!(rb.box && rb.range)) return;
			if (in.mesh_root == kNoIndex) return;
			V3 lo, ld;
			float l;
			ray_to_local(in, wo, wd, lo, ld, l);
			const float lnear = fmul(near_, l), lfar = fmul(far_, l);
			const V3 lrcp = reciprocal_rn(ld);
			const bool lexact = needs_exact(ld);
			const float4 r0 = __ldg(sc.nodes + 2 * size_t(in.mesh_root));
			const float4 r1 = __ldg(sc.nodes + 2 * size_t(in.mesh_root) + 1);
			if (STATS) cnt.mesh_nodes++;
			const SlabResult rr = slab_fast(r0, r1, lo, ld, lrcp, lnear, lfar, lexact);
			if (!(rr.box && rr.range)) return;
			in_mesh = true; mesh_hit = false;
			cur_inst = idx;
			mat_offset = in.mat_offset; mat_count = in.mat_count;
			o = lo; d = ld; rcp = lrcp; exact_only = lexact; len = l;
			sbits = ANY ? 0u : sign_bits(ld);
			near_ = lnear; far_ = lfar;
			cur_begin = __float_as_uint(r1.z);
			cur_tc = __float_as_uint(r1.w);
			state = kTravNode;
		}

		__device__ __forceinline__ bool done() const { return state == kTravDone; }

		// run one ray to completion (used where rays are not pulled dynamically)
		template <bool STATS>
		__device__ __forceinline__ void run(const DScene& sc, Stack& st, TraceCounters& cnt)
		{
			while (!done())
			{
				inner_phase<STATS>(sc, st, cnt);
				leaf_phase<STATS>(sc, st, cnt);
				transition_phase<STATS>(sc, st, cnt);
			}
		}
	};
}
