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
//   * one flat loop of rounds (trav_descend / trav_leaf / trav_pop) serves both levels; the world ray and the committed hit
//     live in shared memory while a mesh is being walked (they are only touched at instance transitions), which keeps the
//     exact closest-hit kernel at 72 registers (7 blocks per SM) and the conservative ones at 64 (8 blocks);
//   * the kernels are bound by instruction issue at low SIMT efficiency, so what few lanes execute is kept short: one
//     straight-line child selection, no out-parameter that forces a value into local memory, the stack addressed as
//     shared memory, no world-ray restore when a walk is over (profiles/r02_final2_phase_lines.txt).
// Scenes with own trees (RZB_SCENE_OWN_TREES) and all shadow queries use a conservative box test instead (FAST).
// What was measured and dropped (B200, 1M-triangle scene): per-lane ray refill (startup = root test +
// instance transform is too expensive to run for single lanes), warp-private work chunks, vote-scheduled phases, a
// camera-ray / bounce-ray queue split, warp-synchronised and while-while rounds, several rays per lane, cache hints --
// none beat free-running lanes on whole-warp batches of 32 rays that the order pass put next to each other.
#pragma once

#include "rzb_device.cuh"

namespace rzb
{
	constexpr float kSlabMargin = 4.0e-7f; // > 2 ulp relative (2^-22 = 2.4e-7)
	constexpr float kInf = __builtin_huge_valf();
	// ---- wide (4-ary) mesh trees of the own-tree mode (RZB_SCENE_WIDE_TREES; built by rzb_set_scene from the uploaded
	// binary trees). A node is 8 x float4: min.x[4], min.y[4], min.z[4], max.x[4], max.y[4], max.z[4], ref[4], spare.
	// ref (30 bits, also what a deferred-node stack entry holds): leaf = bit 29 | count << 25 | first triangle (global,
	// < 2^25); inner = index of the wide node; kWideEmpty = no child.
	constexpr uint32_t kWideLeafBit = 1u << 29;
	constexpr uint32_t kWideEmpty = 0x3FFFFFFFu;
	__device__ __forceinline__ void wide_decode(const uint32_t ref, uint32_t& cur_begin, uint32_t& cur_tc)
	{
		const bool leaf = (ref & kWideLeafBit) != 0u;
		cur_tc = leaf ? (ref >> 25) & 15u : 0u;
		cur_begin = leaf ? ref & 0x1FFFFFFu : ref;
	}

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
	__device__ __forceinline__ bool tiny_component(const V3& d)
	{
		const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
		return (ax != 0.0f && ax < 1.0e-30f) || (ay != 0.0f && ay < 1.0e-30f) || (az != 0.0f && az < 1.0e-30f);
	}
	__device__ __forceinline__ float margin_for(const V3& d) { return tiny_component(d) ? kInf : kSlabMargin; }
	// Trav::sbits = sign bits of the direction (bits 0-2) | kSbitsTiny: the margin is kept as this one bit, not as a register
	constexpr uint32_t kSbitsTiny = 16u; // (bit 3 stays 0: a node of split type 3 never flips)
	__device__ __forceinline__ uint32_t level_bits(const V3& d) { return sign_bits(d) | (tiny_component(d) ? kSbitsTiny : 0u); }
	__device__ __forceinline__ float margin_of(const uint32_t sbits) { return (sbits & kSbitsTiny) ? kInf : kSlabMargin; }

	// Fast slab predicate. Returns true when the box is hit AND its entry distance is within the range; tmin_out is
	// the (approximate, or exact after the fallback) entry distance.
	// CONSERVATIVE = true (scenes whose trees are not the reference's, RZB_SCENE_OWN_TREES): no decision has to equal
	// the reference's, the test only must never reject a box the ray enters -- the interval is widened by 4 ulp
	// instead of being checked against the error bound (10 instructions per box less, no exact fallback).
	template <bool CONSERVATIVE>
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
		if (CONSERVATIVE)
		{
			// a box that matters has tmax > 0; entry distances that matter for the far test are positive
			const float lo = tmin * 0.9999995f, hi = tmax * 1.0000005f;
			tmin_out = lo;
			return !(hi < near_ || lo > hi || lo > far_);
		}
		// closest distance between any two compared quantities versus the error bound of the approximate ones;
		// infinities (a direction component == 0) are exact in both formulations and never "close"
		const float gap = fminf(fminf(fabsf(tmax - near_), fabsf(tmin - tmax)), fabsf(tmin - far_));
		const float bound = margin * fmaxf(fminf(fmaxf(fabsf(tmin), fabsf(tmax)), 1.0e30f), 1.0e-30f);
		tmin_out = tmin;
		if (gap <= bound)
		{
			// (through a temporary: a variable whose address goes to the out-of-line function lives in local memory)
			float te;
			const uint32_t r = slab_exact(n0, n1, o, d, near_, far_, te);
			tmin_out = te;
			return r == 3u;
		}
		return !(tmax < near_ || tmin > tmax || tmin > far_);
	}

	// The two boxes of a sibling pair with ONE error-bound check (decision-exact mode): the approximate entry / exit
	// distances of both boxes are formed first, then the smallest of the six compared gaps is set against the bound of the
	// largest magnitude -- more cautious than checking each box on its own (a few more exact fallbacks, never a wrong
	// decision), and 4 instructions per pair cheaper (ncu: the per-box check was 11 % of k_trace_paths' instructions).
	__device__ __forceinline__ void slab_pair_exact(const float4 a0, const float4 a1, const float4 b0, const float4 b1, const V3& o,
		const V3& d, const V3& rcp, const float near_, const float far_, const bool tiny, bool& ha, bool& hb, float& tma, float& tmb)
	{
		float tmin[2], tmax[2];
#pragma unroll
		for (int k = 0; k < 2; ++k)
		{
			const float4 n0 = k ? b0 : a0, n1 = k ? b1 : a1;
			const float t1 = fmul(fsub(n0.x, o.x), rcp.x), t2 = fmul(fsub(n0.w, o.x), rcp.x);
			const float t3 = fmul(fsub(n0.y, o.y), rcp.y), t4 = fmul(fsub(n1.x, o.y), rcp.y);
			const float t5 = fmul(fsub(n0.z, o.z), rcp.z), t6 = fmul(fsub(n1.y, o.z), rcp.z);
			tmin[k] = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
			tmax[k] = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
		}
		const float gap = fminf(fminf(fminf(fabsf(tmax[0] - near_), fabsf(tmin[0] - tmax[0])), fabsf(tmin[0] - far_)),
			fminf(fminf(fabsf(tmax[1] - near_), fabsf(tmin[1] - tmax[1])), fabsf(tmin[1] - far_)));
		const float mag = fmaxf(fmaxf(fabsf(tmin[0]), fabsf(tmax[0])), fmaxf(fabsf(tmin[1]), fabsf(tmax[1])));
		// `tiny` (a direction component whose reciprocal overflows, margin_for): always the exact path. The floor keeps the
		// bound above the spacing of denormal products; an infinite magnitude gives an infinite bound (exact path: correct, and
		// only rays with a zero direction component get there).
		const float bound = kSlabMargin * fmaxf(mag, 1.0e-30f);
		tma = tmin[0]; tmb = tmin[1];
		if (tiny || gap <= bound)
		{
			// (through temporaries: handing tma / tmb themselves to the out-of-line function put them into local memory -- one
			// STL.64 per pair step and two LDL per deferred child in the profile)
			float ta, tb;
			ha = slab_exact(a0, a1, o, d, near_, far_, ta) == 3u;
			hb = slab_exact(b0, b1, o, d, near_, far_, tb) == 3u;
			tma = ta; tmb = tb;
			return;
		}
		ha = !(tmax[0] < near_ || tmin[0] > tmax[0] || tmin[0] > far_);
		hb = !(tmax[1] < near_ || tmin[1] > tmax[1] || tmin[1] > far_);
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

	// Per-lane traversal state (registers). trav_begin / trav_round / trav_end are the three pieces of one query so that
	// a kernel can either run a whole batch to completion (trace_ray) or replace finished lanes' rays between rounds
	// (k_trace_shadow).
	struct Trav
	{
		V3 o, d, rcp;               // current level (world first)
		float near_, far_, len;
		uint32_t sbits;
		bool in_mesh, mesh_hit, lext, committed_ext, alive;
		uint32_t cur_inst, ltri;
		float lb1, lb2;
		uint32_t mat_offset, mat_count;
		uint32_t cur_begin, cur_tc;
		float4 mask;                // any hit
		uint32_t steps, tris;       // STATS only
	};

	template <bool ANY, bool STATS, bool FAST>
	__device__ __forceinline__ void trav_begin(const DScene& sc, Trav& t, const bool active, const V3 origin, const V3 direction,
		const float near_in, const float far_in, Stack& st, ParkedRay& park, TraceCounters& cnt)
	{
		t.mask = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
		t.steps = t.tris = 0u;
		// no instances: the CPU engine's shadow query answers "occluded" (cpu_engine_kernel.cpp:401), the CUDA one "free"
		if (ANY && sc.instance_count == 0u && (sc.flags & RZB_FLAG_CPU_SEMANTICS)) t.mask = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		t.o = origin; t.d = direction;
		t.rcp = reciprocal_rn(direction);
		t.near_ = near_in; t.far_ = far_in; t.len = 1.0f;
		t.sbits = ANY ? 0u : level_bits(direction);
		t.in_mesh = false; t.mesh_hit = false; t.lext = true;
		t.cur_inst = kNoIndex; t.ltri = kNoIndex;
		t.lb1 = 0.0f; t.lb2 = 0.0f;
		t.mat_offset = 0u; t.mat_count = 0u;
		park.ox = origin.x; park.oy = origin.y; park.oz = origin.z; park.near_ = near_in;
		park.dx = direction.x; park.dy = direction.y; park.dz = direction.z; park.far_ = far_in;
		park.b1 = 0.0f; park.b2 = 0.0f; park.tri = kNoIndex; park.inst = kNoIndex;
		t.committed_ext = true;
		st.sp = 0;

		t.alive = active && sc.instance_count != 0u;
		t.cur_begin = 0u; t.cur_tc = 1u;
		if (t.alive)
		{
			const float4 n0 = __ldg(sc.nodes + 2 * size_t(sc.top_root));
			const float4 n1 = __ldg(sc.nodes + 2 * size_t(sc.top_root) + 1);
			if (STATS) cnt.top_nodes++;
			float tmin;
			t.alive = slab_hit<FAST>(n0, n1, t.o, t.d, t.rcp, t.near_, t.far_, margin_of(t.sbits), tmin);
			t.cur_begin = __float_as_uint(n1.z);
			t.cur_tc = __float_as_uint(n1.w);
		}
	}

	// The three pieces of a round. Invariant: an alive lane has a current node whose box test has passed.
	// ---- descend through inner nodes; returns false when neither child was hit (nothing is current: pop next)
	template <bool ANY, bool STATS, bool FAST, bool WIDE = false>
	__device__ __forceinline__ bool trav_descend(const DScene& sc, Trav& t, Stack& st, TraceCounters& cnt)
	{
		const float4* __restrict__ nodes = sc.nodes;
		bool have_cur = t.alive;
		if (t.alive)
		{
			while ((t.cur_tc & 0x3FFFFFFFu) == 0u)
			{
				if (WIDE && t.in_mesh)
				{
					// one 4-ary step: 7 x LDG.128, four conservative slab tests, nearest hit first, the others deferred
					const float4* q = sc.nodes4 + 8 * size_t(t.cur_begin);
					const float4 mnx = __ldg(q), mny = __ldg(q + 1), mnz = __ldg(q + 2), mxx = __ldg(q + 3), mxy = __ldg(q + 4), mxz = __ldg(q + 5);
					const float4 rf = __ldg(q + 6);
					if (STATS) { cnt.mesh_nodes += 4; t.steps++; }
					float tm[4];
					uint32_t rr[4] = {__float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z), __float_as_uint(rf.w)};
					const float bx0[4] = {mnx.x, mnx.y, mnx.z, mnx.w}, bx1[4] = {mxx.x, mxx.y, mxx.z, mxx.w};
					const float by0[4] = {mny.x, mny.y, mny.z, mny.w}, by1[4] = {mxy.x, mxy.y, mxy.z, mxy.w};
					const float bz0[4] = {mnz.x, mnz.y, mnz.z, mnz.w}, bz1[4] = {mxz.x, mxz.y, mxz.z, mxz.w};
#pragma unroll
					for (int c = 0; c < 4; ++c)
					{
						const float t1 = fmul(fsub(bx0[c], t.o.x), t.rcp.x), t2 = fmul(fsub(bx1[c], t.o.x), t.rcp.x);
						const float t3 = fmul(fsub(by0[c], t.o.y), t.rcp.y), t4 = fmul(fsub(by1[c], t.o.y), t.rcp.y);
						const float t5 = fmul(fsub(bz0[c], t.o.z), t.rcp.z), t6 = fmul(fsub(bz1[c], t.o.z), t.rcp.z);
						const float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
						const float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
						const float lo = tmin * 0.9999995f, hi = tmax * 1.0000005f;
						// (a slab test cannot tell an inverted box from a real one: empty slots are recognised by their reference)
						tm[c] = (rr[c] != kWideEmpty && !(hi < t.near_ || lo > hi || lo > t.far_)) ? lo : kInf;
					}
					if (!ANY)
					{
						// ascending entry distance (5-comparator network); misses (inf) sink to the end
#define RZB_CSWAP(a, b) { const bool sw = tm[b] < tm[a]; const float ta = sw ? tm[b] : tm[a], tb = sw ? tm[a] : tm[b]; \
	const uint32_t ra = sw ? rr[b] : rr[a], rb = sw ? rr[a] : rr[b]; tm[a] = ta; tm[b] = tb; rr[a] = ra; rr[b] = rb; }
						RZB_CSWAP(0, 1) RZB_CSWAP(2, 3) RZB_CSWAP(0, 2) RZB_CSWAP(1, 3) RZB_CSWAP(1, 2)
#undef RZB_CSWAP
						if (!(tm[0] < kInf)) { have_cur = false; break; }
						if (tm[3] < kInf) st.push(kEntryMeshNode | rr[3], __float_as_uint(tm[3]));
						if (tm[2] < kInf) st.push(kEntryMeshNode | rr[2], __float_as_uint(tm[2]));
						if (tm[1] < kInf) st.push(kEntryMeshNode | rr[1], __float_as_uint(tm[1]));
						wide_decode(rr[0], t.cur_begin, t.cur_tc);
					}
					else
					{
						uint32_t first = kWideEmpty;
#pragma unroll
						for (int c = 0; c < 4; ++c)
							if (tm[c] < kInf)
							{
								if (first == kWideEmpty) first = rr[c];
								else st.push(kEntryMeshNode | rr[c], __float_as_uint(tm[c]));
							}
						if (first == kWideEmpty) { have_cur = false; break; }
						wide_decode(first, t.cur_begin, t.cur_tc);
					}
					continue;
				}
				const float4* pair = nodes + 2 * size_t(t.cur_begin); // 64-byte aligned sibling pair
				const float4 p0 = __ldg(pair), p1 = __ldg(pair + 1), p2 = __ldg(pair + 2), p3 = __ldg(pair + 3);
				if (STATS) { if (t.in_mesh) cnt.mesh_nodes += 2; else cnt.top_nodes += 2; t.steps++; }
				float tm0, tm1;
				bool h0, h1;
				if (FAST)
				{
					h0 = slab_hit<true>(p0, p1, t.o, t.d, t.rcp, t.near_, t.far_, 0.0f, tm0);
					h1 = slab_hit<true>(p2, p3, t.o, t.d, t.rcp, t.near_, t.far_, 0.0f, tm1);
				}
				else slab_pair_exact(p0, p1, p2, p3, t.o, t.d, t.rcp, t.near_, t.far_, (t.sbits & kSbitsTiny) != 0u, h0, h1, tm0, tm1);
				// near child first: `flip` = the second child is the near one
				// own trees (FAST) and any hit (its result does not depend on the order; measured: an occluder is found sooner than
				// with the first child first): nearer entry first; reference trees: by ray sign on the split axis, as the reference
				const bool flip = (ANY || FAST) ? (h0 && h1 && tm1 < tm0) : ((t.sbits >> (t.cur_tc >> 30)) & 1u) != 0u;
				// one straight-line selection for "both hit" (near child first, the other deferred) and "one hit" -- as two
				// branches the lanes of a warp ran them one after the other
				if (!(h0 || h1))
				{
					have_cur = false;
					break;
				}
				const bool take_b = h1 && (!h0 || flip); // continue with the second child of the pair
				// the other child is deferred with its entry distance: it is range-tested again when popped, i.e. after the
				// near subtree has been searched
				if (h0 && h1) st.push((t.in_mesh ? kEntryMeshNode : kEntryTopNode) | (t.cur_begin + (take_b ? 0u : 1u)),
					__float_as_uint(take_b ? tm0 : tm1));
				t.cur_tc = __float_as_uint(take_b ? p3.w : p1.w);
				t.cur_begin = __float_as_uint(take_b ? p3.z : p1.z);
			}
		}
		return have_cur;
	}

	// ---- the current node is a leaf: instance-tree leaf = a range of instances to enter one by one; mesh leaf = triangles
	template <bool ANY, bool STATS, bool FAST, bool WIDE = false>
	__device__ __forceinline__ void trav_leaf(const DScene& sc, Trav& t, Stack& st, TraceCounters& cnt)
	{
		{
			const uint32_t count = t.cur_tc & 0x3FFFFFFFu;
			if (!t.in_mesh) st.push(kEntryInstRange | t.cur_begin, t.cur_begin + count);
			else
			{
				const uint32_t end = t.cur_begin + count;
				for (uint32_t i = t.cur_begin; i < end; ++i)
				{
					if (STATS) { cnt.triangles++; t.tris++; }
					if (!ANY)
					{
						if (triangle_closest(sc.tri_hot, i, t.o, t.d, t.near_, t.far_, t.lb1, t.lb2, t.lext))
						{
							t.ltri = i;
							t.mesh_hit = true;
						}
					}
					else
					{
						float tf = t.far_, tb1, tb2;
						bool text;
						if (!triangle_closest(sc.tri_hot, i, t.o, t.d, t.near_, tf, tb1, tb2, text)) continue;
						if (sc.flags & RZB_FLAG_CPU_SEMANTICS) t.mask = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
						else
						{
							const float4 a = shadow_attenuation(sc, i, tb1, tb2, t.mat_offset, t.mat_count);
							t.mask = make_float4(t.mask.x * a.x, t.mask.y * a.y, t.mask.z * a.z, t.mask.w * a.w);
						}
						if (t.mask.w < 1.0e-4f)
						{
							t.alive = false; // occluded
							break;
						}
					}
				}
			}
		}
	}

	// ---- pop until a node with a passed box is current (or the ray is finished)
	template <bool ANY, bool STATS, bool FAST, bool WIDE = false>
	__device__ __forceinline__ void trav_pop(const DScene& sc, Trav& t, Stack& st, ParkedRay& park, TraceCounters& cnt)
	{
		const float4* __restrict__ nodes = sc.nodes;
		while (t.alive)
		{
			const bool have = st.sp != 0;
			uint2 e = make_uint2(kEntryTopNode, 0u);
			if (have) e = st.pop();
			const uint32_t kind = e.x & kEntryKindMask;
			if (t.in_mesh && (!have || kind != kEntryMeshNode))
			{
				// the current mesh is exhausted: leave the instance (cuda_instance.cuh:203-213)
				if (!ANY && t.mesh_hit)
				{
					park.inst = t.cur_inst; park.tri = t.ltri; park.b1 = t.lb1; park.b2 = t.lb2;
					t.committed_ext = t.lext;
					park.near_ = fdiv(t.near_, t.len);
					park.far_ = fdiv(t.far_, t.len);
				}
				// (nothing left on the stack: the walk is over and the world ray is not needed again)
				if (have)
				{
					t.o = v3(park.ox, park.oy, park.oz);
					t.d = v3(park.dx, park.dy, park.dz);
					t.rcp = reciprocal_rn(t.d);
					t.sbits = ANY ? 0u : level_bits(t.d);
					t.near_ = park.near_; t.far_ = park.far_;
					t.len = 1.0f;
				}
				t.in_mesh = false;
			}
			if (!have)
			{
				t.alive = false; // finished
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
				if (!slab_hit<FAST>(n0, n1, t.o, t.d, t.rcp, t.near_, t.far_, margin_of(t.sbits), tmin)) continue;
				if (in.mesh_root == kNoIndex) continue;
				V3 lo, ld;
				float l;
				ray_to_local(in, t.o, t.d, lo, ld, l);
				const float lnear = fmul(t.near_, l), lfar = fmul(t.far_, l);
				const V3 lrcp = reciprocal_rn(ld);
				const uint32_t lbits = ANY ? 0u : level_bits(ld);
				const float4 r0 = __ldg(nodes + 2 * size_t(in.mesh_root));
				const float4 r1 = __ldg(nodes + 2 * size_t(in.mesh_root) + 1);
				if (STATS) cnt.mesh_nodes++;
				if (!slab_hit<FAST>(r0, r1, lo, ld, lrcp, lnear, lfar, margin_of(lbits), tmin)) continue;
				t.in_mesh = true; t.mesh_hit = false;
				t.cur_inst = idx;
				t.mat_offset = in.mat_offset; t.mat_count = in.mat_count;
				t.o = lo; t.d = ld; t.rcp = lrcp; t.len = l;
				t.sbits = lbits;
				t.near_ = lnear; t.far_ = lfar;
				t.cur_begin = __float_as_uint(r1.z);
				t.cur_tc = __float_as_uint(r1.w);
				if (WIDE) wide_decode(__ldg(sc.inst_root4 + idx), t.cur_begin, t.cur_tc);
				break;
			}
			// a deferred node of the current level
			if (!ANY && FAST)
			{
				if (__uint_as_float(e.y) > t.far_) continue; // the stored entry distance is already the conservative one
			}
			else if (!ANY)
			{
				// late range test (the reference tests the far child after the near subtree has been searched)
				const float tmin = __uint_as_float(e.y);
				const float bound = margin_of(t.sbits) * fmaxf(fminf(fabsf(tmin), 1.0e30f), 1.0e-30f);
				if (tmin > t.far_ + bound) continue;
				if (!(tmin < t.far_ - bound))
				{
					// too close to call with the approximate entry distance: evaluate the reference's arithmetic
					const float4 x0 = __ldg(nodes + 2 * size_t(idx));
					const float4 x1 = __ldg(nodes + 2 * size_t(idx) + 1);
					float texact;
					if (!(slab_exact(x0, x1, t.o, t.d, t.near_, t.far_, texact) & 2u)) continue;
				}
			}
			if (WIDE && t.in_mesh)
			{
				wide_decode(idx, t.cur_begin, t.cur_tc); // the entry IS the child reference: no node fetch at pop time
				break;
			}
			const float4 n1 = __ldg(nodes + 2 * size_t(idx) + 1);
			t.cur_begin = __float_as_uint(n1.z);
			t.cur_tc = __float_as_uint(n1.w);
			break;
		}
	}

	// One round: descend from the current node to a leaf, intersect it, pop until a node with a passed box is current
	// (or the ray is finished).
	template <bool ANY, bool STATS, bool FAST, bool WIDE = false>
	__device__ __forceinline__ void trav_round(const DScene& sc, Trav& t, Stack& st, ParkedRay& park, TraceCounters& cnt)
	{
		const bool have_cur = trav_descend<ANY, STATS, FAST, WIDE>(sc, t, st, cnt);
		if (have_cur) trav_leaf<ANY, STATS, FAST, WIDE>(sc, t, st, cnt);
		trav_pop<ANY, STATS, FAST, WIDE>(sc, t, st, park, cnt);
	}

	__device__ __forceinline__ void trav_end(const Trav& t, const bool active, const float near_in, const float far_in,
		const ParkedRay& park, RayResult& res)
	{
		res.t = far_in; res.near_ = near_in; res.b1 = 0.0f; res.b2 = 0.0f;
		res.tri = kNoIndex; res.inst = kNoIndex; res.external = true;
		res.mask = t.mask;
		res.steps = t.steps; res.tris = t.tris;
		if (active)
		{
			res.t = park.far_; res.near_ = park.near_; res.b1 = park.b1; res.b2 = park.b2;
			res.tri = park.tri; res.inst = park.inst; res.external = t.committed_ext;
		}
	}

	// One query per lane, run to completion. All 32 lanes of the warp must call this together; lanes without a ray pass
	// active = false.
	// SYNC = true makes the outer loop warp-uniform (one __any_sync per round): in every round the lanes descend
	// together, then intersect their leaves together, then pop / change level together. SYNC = false lets every lane
	// run its own rounds. Measured on B200 (1M-triangle scene, ms per pass; profiles/): closest hit 1.38 free-running
	// vs 1.79 synchronised. Any hit with the exact box test (80 registers): 0.75 free-running vs 0.43 synchronised;
	// with the conservative test (64 registers, 8 blocks per SM): 0.25 free-running vs 0.35 synchronised -- every
	// render kernel now instantiates SYNC = false, only the one-warp k_raycast keeps SYNC = true.
	// FAST = true selects the conservative box test (slab_hit) and nearer-entry-first child order.
	// MODE: 0 = free-running lanes, 1 = warp-synchronised rounds.
	// Measured and dropped in round 2 ("while-while" rounds): a lane keeps popping and descending until a MESH leaf is current,
	// then the warp's lanes intersect their leaves together (in trav_round a lane whose descent ends without a leaf sits out
	// the leaf phase: triangle tests run at 6.8 of 32 lanes on the 1M-triangle scene). Same records, but 1.44 instead of
	// 0.94 ms per pass (materials scene 0.95 / 0.66; shadow kernel 0.32 / 0.21): every point where lanes wait for each other
	// costs more than the lanes it fills -- like the synchronised rounds and the multi-ray kernels before.
	template <bool ANY, bool STATS, int MODE = (ANY ? 1 : 0), bool FAST = false, bool WIDE = false>
	__device__ __forceinline__ void trace_ray(const DScene& sc, const bool active, const V3 origin, const V3 direction,
		const float near_in, const float far_in, Stack& st, ParkedRay& park, TraceCounters& cnt, RayResult& res)
	{
		Trav t;
		trav_begin<ANY, STATS, FAST>(sc, t, active, origin, direction, near_in, far_in, st, park, cnt);
		while (MODE == 1 ? __any_sync(0xFFFFFFFFu, t.alive) != 0 : t.alive)
			trav_round<ANY, STATS, FAST, WIDE>(sc, t, st, park, cnt);
		trav_end(t, active, near_in, far_in, park, res);
	}
}
