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
//   * sibling pairs are one 64-byte aligned fetch (4 x LDG.128 from one base address); triangles are a 48-byte hot
//     record (3 x LDG.128);
//   * the six IEEE divisions of a slab test are replaced by multiplications with the reciprocal direction and an
//     error margin: the product differs from the correctly rounded quotient by < 2 ulp, so whenever the three
//     comparisons of the predicate are decided by more than that margin the decision equals the reference's;
//     otherwise (about one test in 10^5) the exact divisions are evaluated. ncu on the first version of this
//     kernel showed 38 % of all issued instructions inside division sequences (profiles/);
//   * the far child is deferred on a short stack (shared memory, interleaved by lane) together with its entry
//     distance, so the reference's late range test is a compare at pop time instead of a second box test;
//   * one flat while-while loop serves both levels; the world ray and the committed hit live in shared memory while a
//     mesh is being walked (they are only touched at instance transitions), which keeps the loop at <= 80 registers.
// What was measured and dropped (B200, 1M-triangle scene): per-lane ray refill (startup = root test +
// instance transform is too expensive to run for single lanes), warp-private work chunks, vote-scheduled phases and a
// camera-ray / bounce-ray queue split -- none beat whole-warp batches of 32 neighbouring slots.
#pragma once

#include "rzb_device.cuh"

namespace rzb
{
	constexpr float kSlabMargin = 4.0e-7f; // > 2 ulp relative (2^-22 = 2.4e-7)
	constexpr float kInf = __builtin_huge_valf();

	// exact predicate (the reference's arithmetic); bit 0 = box, bit 1 = range; tmin returned through the reference
	__device__ __noinline__ uint32_t slab_exact(const float4 n0, const float4 n1, const V3 o, const V3 d,
		const float near_, const float far_, float& tmin)
	{
		const bool box = slab_rn(n0, n1, o, d, near_, tmin);
		return (box ? 1u : 0u) | (range_ok(tmin, far_) ? 2u : 0u);
	}

	__device__ __forceinline__ V3 reciprocal_rn(const V3& d) { return v3(fdiv(1.0f, d.x), fdiv(1.0f, d.y), fdiv(1.0f, d.z)); }
	// relative margin of the fast slab test for a ray direction: kSlabMargin, or infinity (= always take the exact
	// path) when a component is so small that its reciprocal overflows while quotients may not
	__device__ __forceinline__ float margin_for(const V3& d)
	{
		const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
		const bool tiny = (ax != 0.0f && ax < 1.0e-30f) || (ay != 0.0f && ay < 1.0e-30f) || (az != 0.0f && az < 1.0e-30f);
		return tiny ? kInf : kSlabMargin;
	}

	// Fast slab predicate. Returns true when the box is hit AND its entry distance is within the range; tmin_out is
	// the (approximate, or exact after the fallback) entry distance.
	__device__ __forceinline__ bool slab_hit(const float4 n0, const float4 n1, const V3& o, const V3& d, const V3& rcp,
		const float near_, const float far_, const float margin, float& tmin_out)
	{
		const float t1 = fmul(fsub(n0.x, o.x), rcp.x);
		const float t2 = fmul(fsub(n0.w, o.x), rcp.x);
		const float t3 = fmul(fsub(n0.y, o.y), rcp.y);
		const float t4 = fmul(fsub(n1.x, o.y), rcp.y);
		const float t5 = fmul(fsub(n0.z, o.z), rcp.z);
		const float t6 = fmul(fsub(n1.y, o.z), rcp.z);
		const float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
		const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
		// closest distance between any two compared quantities versus the error bound of the approximate ones;
		// infinities (a direction component == 0) are exact in both formulations and never "close"
		const float gap = fminf(fminf(fabsf(tmax - near_), fabsf(tmin - tmax)), fabsf(tmin - far_));
		const float bound = margin * fmaxf(fminf(fmaxf(fabsf(tmin), fabsf(tmax)), 1.0e30f), 1.0e-30f);
		tmin_out = tmin;
		if (gap <= bound)
		{
			const uint32_t r = slab_exact(n0, n1, o, d, near_, far_, tmin_out);
			return r == 3u;
		}
		return !(tmax < near_ || tmin > tmax || tmin > far_);
	}

	struct RayResult
	{
		// closest hit
		float t, near_;      // ray.near_far after traversal
		float b1, b2;
		uint32_t tri, inst;  // BVH-order triangle, BVH-order instance (kNoIndex = miss)
		bool external;
		// any hit
		float4 mask;
		// per-ray work (STATS only)
		uint32_t steps, tris;
	};

	// every intersected triangle multiplies the shadow mask by its material's opacity colour (defined in rzb_shade.cuh)
	__device__ __forceinline__ float4 shadow_attenuation(const DScene& sc, const uint32_t tri, const float b1, const float b2,
		const uint32_t mat_offset, const uint32_t mat_count);

	// What a lane parks in shared memory while it walks a mesh (touched only at instance transitions).
	struct __align__(16) ParkedRay
	{
		float ox, oy, oz, near_;
		float dx, dy, dz, far_;
		float b1, b2;
		uint32_t tri, inst;
	};
	static_assert(sizeof(ParkedRay) == 48, "ParkedRay");

	// All 32 lanes of the warp must call this together; lanes without a ray pass active = false.
	// SYNC = true makes the outer loop warp-uniform (one __any_sync per round): in every round the lanes descend
	// together, then intersect their leaves together, then pop / change level together. SYNC = false lets every lane
	// run its own rounds. Measured on B200 (1M-triangle scene, ms per pass; profiles/): closest hit 1.38 free-running
	// vs 1.79 synchronised; any hit 0.75 free-running (triangle code at ~2 of 32 lanes) vs 0.43 synchronised -- so the
	// closest-hit kernels instantiate SYNC = false and the shadow kernels SYNC = true.
	template <bool ANY, bool STATS, bool SYNC = ANY>
	__device__ __forceinline__ void trace_ray(const DScene& sc, const bool active, const V3 origin, const V3 direction,
		const float near_in, const float far_in, Stack& st, ParkedRay& park, TraceCounters& cnt, RayResult& res)
	{
		res.t = far_in; res.near_ = near_in; res.b1 = 0.0f; res.b2 = 0.0f;
		res.tri = kNoIndex; res.inst = kNoIndex; res.external = true;
		res.mask = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
		res.steps = res.tris = 0u;
		// no instances: the CPU engine's shadow query answers "occluded" (cpu_engine_kernel.cpp:401), the CUDA one "free"
		if (ANY && sc.instance_count == 0u && (sc.flags & RZB_FLAG_CPU_SEMANTICS)) res.mask = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		const float4* __restrict__ nodes = sc.nodes;

		// current level (world first)
		V3 o = origin, d = direction;
		V3 rcp = reciprocal_rn(d);
		float margin = margin_for(d);
		float near_ = near_in, far_ = far_in, len = 1.0f;
		uint32_t sbits = ANY ? 0u : sign_bits(d);
		bool in_mesh = false, mesh_hit = false, lext = true;
		uint32_t cur_inst = kNoIndex, ltri = kNoIndex;
		float lb1 = 0.0f, lb2 = 0.0f;
		uint32_t mat_offset = 0u, mat_count = 0u;
		park.ox = origin.x; park.oy = origin.y; park.oz = origin.z; park.near_ = near_in;
		park.dx = direction.x; park.dy = direction.y; park.dz = direction.z; park.far_ = far_in;
		park.b1 = 0.0f; park.b2 = 0.0f; park.tri = kNoIndex; park.inst = kNoIndex;
		bool committed_ext = true;
		st.sp = 0;

		bool alive = active && sc.instance_count != 0u;
		uint32_t cur_begin = 0u, cur_tc = 1u;
		if (alive)
		{
			const float4 n0 = __ldg(nodes + 2 * size_t(sc.top_root));
			const float4 n1 = __ldg(nodes + 2 * size_t(sc.top_root) + 1);
			if (STATS) cnt.top_nodes++;
			float tmin;
			alive = slab_hit(n0, n1, o, d, rcp, near_, far_, margin, tmin);
			cur_begin = __float_as_uint(n1.z);
			cur_tc = __float_as_uint(n1.w);
		}

		while (SYNC ? __any_sync(0xFFFFFFFFu, alive) != 0 : alive)
		{
			// invariant: an alive lane has a current node whose box test has passed
			bool have_cur = alive;
			// ---- descend through inner nodes
			if (alive)
			{
				while ((cur_tc & 0x3FFFFFFFu) == 0u)
				{
					const float4* pair = nodes + 2 * size_t(cur_begin); // 64-byte aligned sibling pair
					const float4 p0 = __ldg(pair), p1 = __ldg(pair + 1), p2 = __ldg(pair + 2), p3 = __ldg(pair + 3);
					if (STATS) { if (in_mesh) cnt.mesh_nodes += 2; else cnt.top_nodes += 2; res.steps++; }
					float tm0, tm1;
					const bool h0 = slab_hit(p0, p1, o, d, rcp, near_, far_, margin, tm0);
					const bool h1 = slab_hit(p2, p3, o, d, rcp, near_, far_, margin, tm1);
					// near child first: `flip` = the second child is the near one
					const bool flip = !ANY && ((sbits >> (cur_tc >> 30)) & 1u) != 0u;
					const bool hit_a = flip ? h1 : h0, hit_b = flip ? h0 : h1;
					if (hit_a)
					{
						// B is deferred with its entry distance: it is range-tested again when popped, i.e. after A's subtree
						if (hit_b) st.push((in_mesh ? kEntryMeshNode : kEntryTopNode) | (cur_begin + (flip ? 0u : 1u)),
							__float_as_uint(flip ? tm0 : tm1));
						cur_tc = __float_as_uint(flip ? p3.w : p1.w);
						cur_begin = __float_as_uint(flip ? p3.z : p1.z);
					}
					else if (hit_b)
					{
						cur_tc = __float_as_uint(flip ? p1.w : p3.w);
						cur_begin = __float_as_uint(flip ? p1.z : p3.z);
					}
					else
					{
						have_cur = false;
						break;
					}
				}
			}
			// ---- leaf
			if (have_cur)
			{
				const uint32_t count = cur_tc & 0x3FFFFFFFu;
				if (!in_mesh) st.push(kEntryInstRange | cur_begin, cur_begin + count);
				else
				{
					const uint32_t end = cur_begin + count;
					for (uint32_t i = cur_begin; i < end; ++i)
					{
						if (STATS) { cnt.triangles++; res.tris++; }
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
							if (sc.flags & RZB_FLAG_CPU_SEMANTICS) res.mask = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
							else
							{
								const float4 a = shadow_attenuation(sc, i, tb1, tb2, mat_offset, mat_count);
								res.mask = make_float4(res.mask.x * a.x, res.mask.y * a.y, res.mask.z * a.z, res.mask.w * a.w);
							}
							if (res.mask.w < 1.0e-4f)
							{
								alive = false; // occluded
								break;
							}
						}
					}
				}
			}
			// ---- pop until a node with a passed box is current (or the ray is finished)
			while (alive)
			{
				const bool have = st.sp != 0;
				uint2 e = make_uint2(kEntryTopNode, 0u);
				if (have) e = st.pop();
				const uint32_t kind = e.x & kEntryKindMask;
				if (in_mesh && (!have || kind != kEntryMeshNode))
				{
					// the current mesh is exhausted: leave the instance (cuda_instance.cuh:203-213)
					if (!ANY && mesh_hit)
					{
						park.inst = cur_inst; park.tri = ltri; park.b1 = lb1; park.b2 = lb2;
						committed_ext = lext;
						park.near_ = fdiv(near_, len);
						park.far_ = fdiv(far_, len);
					}
					o = v3(park.ox, park.oy, park.oz);
					d = v3(park.dx, park.dy, park.dz);
					rcp = reciprocal_rn(d);
					margin = margin_for(d);
					sbits = ANY ? 0u : sign_bits(d);
					near_ = park.near_; far_ = park.far_;
					len = 1.0f;
					in_mesh = false;
				}
				if (!have)
				{
					alive = false; // finished
					break;
				}
				const uint32_t idx = e.x & kEntryIndexMask;
				if (kind == kEntryInstRange)
				{
					const uint32_t end = e.y;
					if (idx + 1u < end) st.push(kEntryInstRange | (idx + 1u), end);
					// Instance::closestIntersection / anyIntersection (cuda_instance.cuh:186-229)
					if (STATS) cnt.instances++;
					const DInstance in = load_instance(sc.instances, idx);
					const float4 n0 = make_float4(in.bminx, in.bminy, in.bminz, in.bmaxx);
					const float4 n1 = make_float4(in.bmaxy, in.bmaxz, 0.0f, 0.0f);
					float tmin;
					if (!slab_hit(n0, n1, o, d, rcp, near_, far_, margin, tmin)) continue;
					if (in.mesh_root == kNoIndex) continue;
					V3 lo, ld;
					float l;
					ray_to_local(in, o, d, lo, ld, l);
					const float lnear = fmul(near_, l), lfar = fmul(far_, l);
					const V3 lrcp = reciprocal_rn(ld);
					const float lmargin = margin_for(ld);
					const float4 r0 = __ldg(nodes + 2 * size_t(in.mesh_root));
					const float4 r1 = __ldg(nodes + 2 * size_t(in.mesh_root) + 1);
					if (STATS) cnt.mesh_nodes++;
					if (!slab_hit(r0, r1, lo, ld, lrcp, lnear, lfar, lmargin, tmin)) continue;
					in_mesh = true; mesh_hit = false;
					cur_inst = idx;
					mat_offset = in.mat_offset; mat_count = in.mat_count;
					o = lo; d = ld; rcp = lrcp; margin = lmargin; len = l;
					sbits = ANY ? 0u : sign_bits(ld);
					near_ = lnear; far_ = lfar;
					cur_begin = __float_as_uint(r1.z);
					cur_tc = __float_as_uint(r1.w);
					break;
				}
				// a deferred node of the current level
				if (!ANY)
				{
					// late range test (the reference tests the far child after the near subtree has been searched)
					const float tmin = __uint_as_float(e.y);
					const float bound = margin * fmaxf(fminf(fabsf(tmin), 1.0e30f), 1.0e-30f);
					if (tmin > far_ + bound) continue;
					if (!(tmin < far_ - bound))
					{
						// too close to call with the approximate entry distance: evaluate the reference's arithmetic
						const float4 x0 = __ldg(nodes + 2 * size_t(idx));
						const float4 x1 = __ldg(nodes + 2 * size_t(idx) + 1);
						float texact;
						if (!(slab_exact(x0, x1, o, d, near_, far_, texact) & 2u)) continue;
					}
				}
				const float4 n1 = __ldg(nodes + 2 * size_t(idx) + 1);
				cur_begin = __float_as_uint(n1.z);
				cur_tc = __float_as_uint(n1.w);
				break;
			}
		}
		if (active)
		{
			res.t = park.far_; res.near_ = park.near_; res.b1 = park.b1; res.b2 = park.b2;
			res.tri = park.tri; res.inst = park.inst; res.external = committed_ext;
		}
	}
}
