// Closest-hit traversal with SEVERAL RAYS PER LANE and phase voting ("mr" kernels).
//
// Why: with one ray per lane (rzb_traverse.cuh) a warp's lanes spread over the phases of the walk -- pair step, leaf,
// pop / instance transition, finished -- and every instruction runs for the 8-12 lanes that happen to be in that phase
// (ncu, profiles/: 8.1-11.0 of 32 threads per instruction on the 1M-triangle scene). Here every lane owns kMrRays rays
// whose state lives in shared memory ([field][ray][lane]: conflict-free whichever ray a lane picks). In every round the
// warp votes for the phase most lanes can serve, each lane picks one of ITS rays that is in that phase, and the phase
// runs once for (nearly) all lanes:
//     F  fetch      write the finished ray's record, pull the next slot (one atomic per warp and round), root test
//     N  node       ONE sibling-pair step (64-byte fetch, two slab tests, near child first, far child deferred)
//     T  triangles  the current leaf
//     H  heavy      everything rare: leave a mesh, next instance of a range, instance transform + mesh root test,
//                   end of the walk (reuses the pop loop of rzb_traverse.cuh on a register copy of the state)
// Pops of deferred nodes of the current level are done in line at the end of N and T (a compare and one 16-byte load).
// Each ray still performs exactly the operation sequence of the one-ray-per-lane walk (its own stack, near child first,
// late range test at pop time, the same __f*_rn arithmetic), so every accept / reject decision and every hit record is
// the same; only WHICH lane-round executes a step changes (tests: the ray-set entry point runs on these kernels too).
//
// State per ray in shared memory: 4 x float4 (64 B): {o.xyz, far} {rcp.xyz, near} {cur_begin, cur_tc, flags, ltri}
// {d.xyz, len}; what only instance transitions touch (parked world range and committed hit, barycentrics of the best
// triangle of the current mesh, instance id, slot) and the deferred-node stacks are per-thread local memory.
#pragma once

#include "rzb_traverse.cuh"

namespace rzb
{
	constexpr int kMrMaxRays = 4;  // rays per lane: template parameter K of everything below (2..4)
	constexpr int kMrBlock = 128;  // threads per block
	constexpr int kMrStack = kSmemStack + kLocalStack; // entries per ray (the depth bound rzb_set_scene enforces)

	enum : uint32_t
	{
		kMrEmpty = 0u,  // no ray: wants a fetch
		kMrNode = 1u,   // current node is an inner node of the current level
		kMrLeaf = 2u,   // current node is a triangle leaf
		kMrHeavy = 3u,  // the next step is rare work (see above)
		kMrDone = 4u,   // finished: record waits to be written (served by F)
		kMrDead = 5u    // no ray and no work left
	};
	enum : uint32_t
	{
		kMrInMesh = 1u, kMrMeshHit = 2u, kMrLext = 4u, kMrCommittedExt = 8u, kMrMarginInf = 16u, kMrSbitsShift = 5u
	};

	// what the thread keeps for its kMrRays rays outside shared memory (dynamic index -> local memory; touched rarely)
	template <int K>
	struct MrCold
	{
		float park_near[K], park_far[K], park_b1[K], park_b2[K];
		uint32_t park_tri[K], park_inst[K];
		float lb1[K], lb2[K];
		uint32_t cur_inst[K], handle[K], user[K];
		float wo[K][3], wd[K][3]; // the world ray (restored when a mesh is left)
		uint2 stack[K][kMrStack];
	};

	template <int K>
	struct MrHot
	{
		float4* base; // &smem[0][0][threadIdx.x]
		__device__ __forceinline__ float4& q(const int field, const int k) const { return base[(field * K + k) * kMrBlock]; }
	};
	enum { kQA = 0, kQB = 1, kQC = 2, kQD = 3, kMrFields = 4 };

	struct MrStackView
	{
		uint2* e;
		int sp;
		__device__ __forceinline__ void push(uint32_t a, uint32_t b) { e[sp < kMrStack ? sp : kMrStack - 1] = make_uint2(a, b); ++sp; }
		__device__ __forceinline__ uint2 pop() { --sp; return e[sp < kMrStack ? sp : kMrStack - 1]; }
		__device__ __forceinline__ uint2 peek() const { return e[sp - 1 < kMrStack ? sp - 1 : kMrStack - 1]; }
	};

	// per-lane bookkeeping in registers: one bit mask per phase (bit k = ray k is in that phase; packed 4 bits per phase:
	// phase t occupies bits [4t, 4t+4) of `tags`) and the 8-bit stack pointer of each ray
	struct MrLane
	{
		uint32_t tags, sps, rr; // rr: rotating start of the search, so that a lane's rays take turns
		__device__ __forceinline__ uint32_t mask(const uint32_t t) const { return (tags >> (4u * t)) & 15u; }
		__device__ __forceinline__ void init(const int rays)
		{
			tags = ((1u << rays) - 1u) << (4u * kMrEmpty);
			sps = 0u; rr = 0u;
		}
		// move ray k from phase `from` to phase `to`
		__device__ __forceinline__ void move(const int k, const uint32_t from, const uint32_t to)
		{
			tags = (tags & ~(1u << (4u * from + k))) | (1u << (4u * to + k));
		}
		__device__ __forceinline__ int sp(const int k) const { return int((sps >> (8 * k)) & 255u); }
		__device__ __forceinline__ void set_sp(const int k, const int v) { sps = (sps & ~(255u << (8 * k))) | (uint32_t(v) << (8 * k)); }
		// a ray of this lane in phase t (round-robin), or -1
		__device__ __forceinline__ int pick(const uint32_t t) const
		{
			const uint32_t m = mask(t);
			if (m == 0u) return -1;
			const uint32_t rot = ((m | (m << 4)) >> rr) & 15u; // bit j = ray (rr + j) % 4
			return int((rr + uint32_t(__ffs(rot)) - 1u) & 3u);
		}
	};

	__device__ __forceinline__ void mr_prefetch(const void* p)
	{
#ifndef RZB_HOST_SIM
		asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
		(void)p;
#endif
	}

	// ---- pop of deferred nodes of the CURRENT level (the common case). Returns the ray's next phase; anything that is
	// not a plain deferred node (instance range, end of a mesh, end of the walk) is left to the heavy phase.
	template <bool FAST, bool STATS>
	__device__ __forceinline__ uint32_t mr_pop_inline(const DScene& sc, MrStackView& st, const bool in_mesh, const V3& o, const V3& d,
		const float near_, const float far_, const float margin, uint32_t& cur_begin, uint32_t& cur_tc)
	{
		const float4* __restrict__ nodes = sc.nodes;
		for (;;)
		{
			if (st.sp == 0) return kMrHeavy;
			const uint2 e = st.peek();
			if ((e.x & kEntryKindMask) != (in_mesh ? kEntryMeshNode : kEntryTopNode)) return kMrHeavy;
			st.sp--;
			const uint32_t idx = e.x & kEntryIndexMask;
			if (FAST)
			{
				if (__uint_as_float(e.y) > far_) continue; // the stored entry distance is already the conservative one
			}
			else
			{
				// late range test (the reference tests the far child after the near subtree has been searched)
				const float tmin = __uint_as_float(e.y);
				const float bound = margin * fmaxf(fminf(fabsf(tmin), 1.0e30f), 1.0e-30f);
				if (tmin > far_ + bound) continue;
				if (!(tmin < far_ - bound))
				{
					const float4 x0 = __ldg(nodes + 2 * size_t(idx));
					const float4 x1 = __ldg(nodes + 2 * size_t(idx) + 1);
					float texact;
					if (!(slab_exact(x0, x1, o, d, near_, far_, texact) & 2u)) continue;
				}
			}
			const float4 n1 = __ldg(nodes + 2 * size_t(idx) + 1);
			cur_begin = __float_as_uint(n1.z);
			cur_tc = __float_as_uint(n1.w);
			if ((cur_tc & 0x3FFFFFFFu) == 0u) return kMrNode;
			if (in_mesh) return kMrLeaf;
			// a leaf of the instance tree: its instances are entered one after the other by the heavy phase
			st.push(kEntryInstRange | cur_begin, cur_begin + (cur_tc & 0x3FFFFFFFu));
			return kMrHeavy;
		}
	}

	// what the ray will read first in its next round goes to L1 now (the lane serves its other rays in between)
	__device__ __forceinline__ void mr_prefetch_next(const DScene& sc, const uint32_t next, const uint32_t cur_begin)
	{
		if (next == kMrNode) mr_prefetch(sc.nodes + 2 * size_t(cur_begin));
		else if (next == kMrLeaf) mr_prefetch(sc.tri_hot + 3 * size_t(cur_begin));
	}

	// ---- N: up to STEPS sibling-pair steps while the ray stays on inner nodes
	template <int K, bool FAST, bool STATS, int STEPS>
	__device__ __forceinline__ void mr_node(const DScene& sc, const MrHot<K>& hot, MrCold<K>& cold, MrLane& lane, const int k, TraceCounters& cnt)
	{
		const float4 A = hot.q(kQA, k), B = hot.q(kQB, k), C = hot.q(kQC, k), D = hot.q(kQD, k);
		const V3 o = v3(A.x, A.y, A.z), rcp = v3(B.x, B.y, B.z), d = v3(D.x, D.y, D.z);
		const float far_ = A.w, near_ = B.w;
		uint32_t cur_begin = __float_as_uint(C.x), cur_tc = __float_as_uint(C.y);
		const uint32_t flags = __float_as_uint(C.z);
		const bool in_mesh = (flags & kMrInMesh) != 0u;
		const float margin = (flags & kMrMarginInf) ? kInf : kSlabMargin;
		const uint32_t sbits = (flags >> kMrSbitsShift) & 7u;
		MrStackView st{cold.stack[k], lane.sp(k)};
		uint32_t next = kMrNode;
#pragma unroll 1
		for (int step = 0; step < STEPS && next == kMrNode; ++step)
		{
			const float4* pair = sc.nodes + 2 * size_t(cur_begin); // 64-byte aligned sibling pair
			const float4 p0 = __ldg(pair), p1 = __ldg(pair + 1), p2 = __ldg(pair + 2), p3 = __ldg(pair + 3);
			if (STATS) { if (in_mesh) cnt.mesh_nodes += 2; else cnt.top_nodes += 2; }
			float tm0, tm1;
			const bool h0 = slab_hit<FAST>(p0, p1, o, d, rcp, near_, far_, margin, tm0);
			const bool h1 = slab_hit<FAST>(p2, p3, o, d, rcp, near_, far_, margin, tm1);
			const bool flip = FAST ? (h0 && h1 && tm1 < tm0) : ((sbits >> (cur_tc >> 30)) & 1u) != 0u;
			const bool hit_a = flip ? h1 : h0, hit_b = flip ? h0 : h1;
			if (hit_a || hit_b)
			{
				if (hit_a && hit_b)
					st.push((in_mesh ? kEntryMeshNode : kEntryTopNode) | (cur_begin + (flip ? 0u : 1u)), __float_as_uint(flip ? tm0 : tm1));
				const bool take_second = hit_a ? flip : !flip;
				cur_tc = __float_as_uint(take_second ? p3.w : p1.w);
				cur_begin = __float_as_uint(take_second ? p3.z : p1.z);
				if ((cur_tc & 0x3FFFFFFFu) == 0u) next = kMrNode;
				else if (in_mesh) next = kMrLeaf;
				else
				{
					st.push(kEntryInstRange | cur_begin, cur_begin + (cur_tc & 0x3FFFFFFFu));
					next = kMrHeavy;
				}
			}
			else next = mr_pop_inline<FAST, STATS>(sc, st, in_mesh, o, d, near_, far_, margin, cur_begin, cur_tc);
		}
		mr_prefetch_next(sc, next, cur_begin);
		float2* c2 = reinterpret_cast<float2*>(&hot.q(kQC, k));
		*c2 = make_float2(__uint_as_float(cur_begin), __uint_as_float(cur_tc));
		lane.set_sp(k, st.sp);
		if (next != kMrNode) lane.move(k, kMrNode, next);
	}

	// ---- T: the current leaf
	template <int K, bool FAST, bool STATS>
	__device__ __forceinline__ void mr_leaf(const DScene& sc, const MrHot<K>& hot, MrCold<K>& cold, MrLane& lane, const int k, TraceCounters& cnt)
	{
		const float4 A = hot.q(kQA, k), C = hot.q(kQC, k), D = hot.q(kQD, k);
		const V3 o = v3(A.x, A.y, A.z), d = v3(D.x, D.y, D.z);
		float far_ = A.w;
		const float near_ = hot.q(kQB, k).w;
		uint32_t cur_begin = __float_as_uint(C.x), cur_tc = __float_as_uint(C.y);
		uint32_t flags = __float_as_uint(C.z), ltri = __float_as_uint(C.w);
		float lb1 = 0.0f, lb2 = 0.0f;
		bool lext = (flags & kMrLext) != 0u, hit = false;
		const uint32_t end = cur_begin + (cur_tc & 0x3FFFFFFFu);
		for (uint32_t i = cur_begin; i < end; ++i)
		{
			if (STATS) cnt.triangles++;
			if (triangle_closest(sc.tri_hot, i, o, d, near_, far_, lb1, lb2, lext))
			{
				ltri = i;
				hit = true;
			}
		}
		if (hit)
		{
			cold.lb1[k] = lb1; cold.lb2[k] = lb2;
			flags = (flags & ~uint32_t(kMrLext)) | kMrMeshHit | (lext ? kMrLext : 0u);
			hot.q(kQA, k).w = far_;
		}
		MrStackView st{cold.stack[k], lane.sp(k)};
		const float margin = (flags & kMrMarginInf) ? kInf : kSlabMargin;
		const uint32_t next = mr_pop_inline<FAST, STATS>(sc, st, true, o, d, near_, far_, margin, cur_begin, cur_tc);
		mr_prefetch_next(sc, next, cur_begin);
		hot.q(kQC, k) = make_float4(__uint_as_float(cur_begin), __uint_as_float(cur_tc), __uint_as_float(flags), __uint_as_float(ltri));
		lane.set_sp(k, st.sp);
		if (next != kMrLeaf) lane.move(k, kMrLeaf, next);
	}

	// ---- H: rare work. The pop loop of rzb_traverse.cuh (trav_round) on the ray's full state: leave the mesh (commit
	// its hit to the parked world range), next instance of a range (world box test, G2L transform with IEEE divisions,
	// mesh root test), deferred nodes of the level that was re-entered, end of the walk.
	template <int K, bool FAST, bool STATS>
	__device__ __forceinline__ void mr_heavy(const DScene& sc, const MrHot<K>& hot, MrCold<K>& cold, MrLane& lane, const int k, TraceCounters& cnt)
	{
		const float4* __restrict__ nodes = sc.nodes;
		const float4 A = hot.q(kQA, k), B = hot.q(kQB, k), C = hot.q(kQC, k), D = hot.q(kQD, k);
		V3 o = v3(A.x, A.y, A.z), rcp = v3(B.x, B.y, B.z), d = v3(D.x, D.y, D.z);
		float far_ = A.w, near_ = B.w, len = D.w;
		uint32_t cur_begin = __float_as_uint(C.x), cur_tc = __float_as_uint(C.y);
		uint32_t flags = __float_as_uint(C.z), ltri = __float_as_uint(C.w);
		bool in_mesh = (flags & kMrInMesh) != 0u;
		float margin = (flags & kMrMarginInf) ? kInf : kSlabMargin;
		MrStackView st{cold.stack[k], lane.sp(k)};
		uint32_t next = kMrDone;
		for (;;)
		{
			const bool have = st.sp != 0;
			uint2 e = make_uint2(kEntryTopNode, 0u);
			if (have) e = st.pop();
			const uint32_t kind = e.x & kEntryKindMask;
			if (in_mesh && (!have || kind != kEntryMeshNode))
			{
				// the current mesh is exhausted: leave the instance (cuda_instance.cuh:203-213)
				if (flags & kMrMeshHit)
				{
					cold.park_inst[k] = cold.cur_inst[k]; cold.park_tri[k] = ltri;
					cold.park_b1[k] = cold.lb1[k]; cold.park_b2[k] = cold.lb2[k];
					flags = (flags & ~uint32_t(kMrCommittedExt)) | ((flags & kMrLext) ? kMrCommittedExt : 0u);
					cold.park_near[k] = fdiv(near_, len);
					cold.park_far[k] = fdiv(far_, len);
				}
				o = v3(cold.wo[k][0], cold.wo[k][1], cold.wo[k][2]);
				d = v3(cold.wd[k][0], cold.wd[k][1], cold.wd[k][2]);
				rcp = reciprocal_rn(d);
				margin = margin_for(d);
				near_ = cold.park_near[k]; far_ = cold.park_far[k];
				len = 1.0f;
				in_mesh = false;
				flags = (flags & (kMrCommittedExt | kMrLext)) | (margin == kInf ? kMrMarginInf : 0u) | (sign_bits(d) << kMrSbitsShift);
			}
			if (!have) { next = kMrDone; break; }
			const uint32_t idx = e.x & kEntryIndexMask;
			if (kind == kEntryInstRange)
			{
				const uint32_t end = e.y;
				if (idx + 1u < end) st.push(kEntryInstRange | (idx + 1u), end);
				// Instance::closestIntersection (cuda_instance.cuh:186-214)
				if (STATS) cnt.instances++;
				const DInstance in = load_instance(sc.instances, idx);
				const float4 n0 = make_float4(in.bminx, in.bminy, in.bminz, in.bmaxx);
				const float4 n1 = make_float4(in.bmaxy, in.bmaxz, 0.0f, 0.0f);
				float tmin;
				if (!slab_hit<FAST>(n0, n1, o, d, rcp, near_, far_, margin, tmin)) continue;
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
				if (!slab_hit<FAST>(r0, r1, lo, ld, lrcp, lnear, lfar, lmargin, tmin)) continue;
				in_mesh = true;
				cold.cur_inst[k] = idx;
				o = lo; d = ld; rcp = lrcp; margin = lmargin; len = l;
				near_ = lnear; far_ = lfar;
				flags = (flags & kMrCommittedExt) | kMrInMesh | kMrLext | (margin == kInf ? kMrMarginInf : 0u) | (sign_bits(ld) << kMrSbitsShift);
				ltri = kNoIndex;
				cur_begin = __float_as_uint(r1.z);
				cur_tc = __float_as_uint(r1.w);
				next = (cur_tc & 0x3FFFFFFFu) == 0u ? kMrNode : kMrLeaf;
				break;
			}
			// a deferred node of the (re-entered) current level
			if (FAST)
			{
				if (__uint_as_float(e.y) > far_) continue;
			}
			else
			{
				const float tmin = __uint_as_float(e.y);
				const float bound = margin * fmaxf(fminf(fabsf(tmin), 1.0e30f), 1.0e-30f);
				if (tmin > far_ + bound) continue;
				if (!(tmin < far_ - bound))
				{
					const float4 x0 = __ldg(nodes + 2 * size_t(idx));
					const float4 x1 = __ldg(nodes + 2 * size_t(idx) + 1);
					float texact;
					if (!(slab_exact(x0, x1, o, d, near_, far_, texact) & 2u)) continue;
				}
			}
			const float4 n1 = __ldg(nodes + 2 * size_t(idx) + 1);
			cur_begin = __float_as_uint(n1.z);
			cur_tc = __float_as_uint(n1.w);
			if ((cur_tc & 0x3FFFFFFFu) == 0u) { next = kMrNode; break; }
			if (in_mesh) { next = kMrLeaf; break; }
			st.push(kEntryInstRange | cur_begin, cur_begin + (cur_tc & 0x3FFFFFFFu));
		}
		mr_prefetch_next(sc, next, cur_begin);
		hot.q(kQA, k) = make_float4(o.x, o.y, o.z, far_);
		hot.q(kQB, k) = make_float4(rcp.x, rcp.y, rcp.z, near_);
		hot.q(kQC, k) = make_float4(__uint_as_float(cur_begin), __uint_as_float(cur_tc), __uint_as_float(flags), __uint_as_float(ltri));
		hot.q(kQD, k) = make_float4(d.x, d.y, d.z, len);
		lane.set_sp(k, st.sp);
		if (next != kMrHeavy) lane.move(k, kMrHeavy, next);
	}

	// ---- start of a walk (trav_begin of rzb_traverse.cuh): world ray into the state, root test of the instance tree.
	// The ray is in phase kMrEmpty when this is called.
	template <int K, bool FAST, bool STATS>
	__device__ __forceinline__ void mr_begin(const DScene& sc, const MrHot<K>& hot, MrCold<K>& cold, MrLane& lane, const int k,
		const V3 o, const V3 d, const float near_in, const float far_in, TraceCounters& cnt)
	{
		const V3 rcp = reciprocal_rn(d);
		const float margin = margin_for(d);
		cold.wo[k][0] = o.x; cold.wo[k][1] = o.y; cold.wo[k][2] = o.z;
		cold.wd[k][0] = d.x; cold.wd[k][1] = d.y; cold.wd[k][2] = d.z;
		cold.park_near[k] = near_in; cold.park_far[k] = far_in;
		cold.park_b1[k] = 0.0f; cold.park_b2[k] = 0.0f; cold.park_tri[k] = kNoIndex; cold.park_inst[k] = kNoIndex;
		cold.lb1[k] = 0.0f; cold.lb2[k] = 0.0f; cold.cur_inst[k] = kNoIndex;
		uint32_t flags = kMrCommittedExt | kMrLext | (margin == kInf ? kMrMarginInf : 0u) | (sign_bits(d) << kMrSbitsShift);
		uint32_t cur_begin = 0u, cur_tc = 1u, next = kMrDone;
		MrStackView st{cold.stack[k], 0};
		if (sc.instance_count != 0u)
		{
			const float4 n0 = __ldg(sc.nodes + 2 * size_t(sc.top_root));
			const float4 n1 = __ldg(sc.nodes + 2 * size_t(sc.top_root) + 1);
			if (STATS) cnt.top_nodes++;
			float tmin;
			if (slab_hit<FAST>(n0, n1, o, d, rcp, near_in, far_in, margin, tmin))
			{
				cur_begin = __float_as_uint(n1.z);
				cur_tc = __float_as_uint(n1.w);
				if ((cur_tc & 0x3FFFFFFFu) == 0u) next = kMrNode;
				else
				{
					st.push(kEntryInstRange | cur_begin, cur_begin + (cur_tc & 0x3FFFFFFFu));
					next = kMrHeavy;
				}
			}
		}
		hot.q(kQA, k) = make_float4(o.x, o.y, o.z, far_in);
		hot.q(kQB, k) = make_float4(rcp.x, rcp.y, rcp.z, near_in);
		hot.q(kQC, k) = make_float4(__uint_as_float(cur_begin), __uint_as_float(cur_tc), __uint_as_float(flags), __uint_as_float(kNoIndex));
		hot.q(kQD, k) = make_float4(d.x, d.y, d.z, 1.0f);
		lane.set_sp(k, st.sp);
		lane.move(k, kMrEmpty, next);
	}

	// the finished ray's result (trav_end)
	template <int K>
	__device__ __forceinline__ void mr_result(const MrHot<K>& hot, const MrCold<K>& cold, const int k, RayResult& res)
	{
		const uint32_t flags = __float_as_uint(hot.q(kQC, k).z);
		res.t = cold.park_far[k]; res.near_ = cold.park_near[k]; res.b1 = cold.park_b1[k]; res.b2 = cold.park_b2[k];
		res.tri = cold.park_tri[k]; res.inst = cold.park_inst[k]; res.external = (flags & kMrCommittedExt) != 0u;
		res.mask = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
		res.steps = 0u; res.tris = 0u;
	}

	// Which phase the warp runs this round: the one most lanes can serve (ties: node, leaf, heavy, fetch).
	__device__ __forceinline__ uint32_t mr_vote(const MrLane& lane, const bool work_left)
	{
		const uint32_t fetchable = lane.mask(kMrDone) | (work_left ? lane.mask(kMrEmpty) : 0u);
		const int cn = __popc(__ballot_sync(0xFFFFFFFFu, lane.mask(kMrNode) != 0u));
		const int cl = __popc(__ballot_sync(0xFFFFFFFFu, lane.mask(kMrLeaf) != 0u));
		const int ch = __popc(__ballot_sync(0xFFFFFFFFu, lane.mask(kMrHeavy) != 0u));
		const int cf = __popc(__ballot_sync(0xFFFFFFFFu, fetchable != 0u));
		if (cn >= cl && cn >= ch && cn >= cf && cn > 0) return kMrNode;
		if (cl >= ch && cl >= cf && cl > 0) return kMrLeaf;
		if (ch >= cf && ch > 0) return kMrHeavy;
		if (cf > 0) return kMrEmpty;
		return kMrDead;
	}
}
