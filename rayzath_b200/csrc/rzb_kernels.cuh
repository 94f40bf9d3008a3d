// The wavefront kernels of the B200 render path (one pass = one path segment per pixel, as in the
// reference: cuda_render_kernel.cu:67-121, but split by stage instead of one megakernel):
//
//   k_reset          zero the accumulator, pixel-centre rays into the path state      (passReset + generateCameraRay)
//   k_trace_paths    closest hit for every path; persistent warps pull 32-slot batches with one atomic
//   k_shade          surface analysis, emission, BSDF sampling, NEE set-up (shadow rays appended to a queue by
//                    warp-ballot compaction), accumulation, continue-or-regenerate
//   k_trace_shadow   any-hit over the compacted shadow queue; visible light is added to the accumulator
//   k_tonemap        ComputeFinalColor (cuda_postprocess_kernel.cu:38-58), optionally summing peer accumulators
//                    over NVLink loads in the same pass
//   k_raycast        the pick ray (rayCast, cuda_render_kernel.cu:130-144)
//   k_pack_*         scene repacking on the device (rzb_set_scene)
// plus the ray-set entry points used for ID parity (k_trace_rays, k_trace_any_rays, k_convert_hits,
// k_camera_rays).
//
// Path state lives in HBM as three coalesced arrays indexed by slot (40 B per pixel): 256 consecutive slots cover a
// 16x16-pixel chunk (eight 8x4 tiles), so the 32 rays of a batch start in one 8x4 screen patch (better node and
// triangle reuse in L1/L2 than the reference's 32x1 rows) and a context can own a row band or interleaved chunk rows.
#pragma once

#include "rzb_shade.cuh"
#include "rzb_traverse_mr.cuh"

namespace rzb
{
	constexpr uint32_t kHitTriMask = 0x3FFFFFFFu;
	constexpr uint32_t kHitExternalBit = 0x80000000u;
	constexpr uint32_t kHitScatterBit = 0x40000000u;
	constexpr uint32_t kMediumShift = 8u;

	struct DFrame
	{
		DCamera cam;
		uint32_t tiles_x, tiles_y, n_slots;
		uint64_t tiles_x_magic;         // ceil(2^40 / tiles_x), see slot_to_pixel
		uint32_t row_begin, row_end;    // tile split: this context renders image rows [row_begin, row_end)
		uint32_t slot_begin, slot_end;  // the 256-slot chunks that cover those rows
		uint32_t il_index, il_count;    // interleaved tile split: only 16-row chunk rows r with r % il_count == il_index
		float4* st_o;   // {o.xyz, bits(depth | medium << 8)}
		float4* st_d;   // {d.xyz, throughput.r}
		float2* st_c;   // {throughput.g, throughput.b}
		float4* hit_a;  // {t, b1, b2, bits(tri | flags)}
		uint32_t* hit_inst;
		float4* accum;  // row-major width*height: rgb sum, alpha = completed paths
		float* depth;   // row-major, written on the first pass after a reset
		// shadow queue
		float4* sh_o;   // {origin.xyz, max distance}
		float4* sh_d;   // {direction.xyz, bits(pixel index)}
		float4* sh_c;   // {contribution.rgb, -}
		uint32_t* counters; // [0] closest work counter, [1] shadow queue size, [2] shadow work counter
		unsigned long long* work; // RZB_FLAG_COUNT_WORK: [0..3] closest top/inst/mesh/tri, [4..7] shadow, [8] shadow rays
		uint32_t shadow_capacity;
		uint32_t pass_index;
		// temporal reprojection (RZB_FLAG_TEMPORAL_REPROJECTION): the frame that the current one replaces
		const float4* prev_accum;
		const float* prev_depth;
		DCamera prev_cam;
		float reproject_blend; // 0 = off

		uint32_t max_depth, direct_samples, spot_samples;
		float inv_pdf_direct, inv_pdf_spot; // light count / samples
		uint64_t seed;
		// ray ordering: k_shade bins the next pass's rays and the shadow rays it queues (one atomic each: bin counter ->
		// rank inside the bin); after a prefix sum over the bins k_scatter_order writes slot / queue indices in bin
		// order; the traversal kernels pull their batches through `order` / `sh_order` (NULL = slot / append order)
		uint32_t* sort_keys;      // [slot - slot_begin] bin of the slot's next ray
		uint32_t* sort_rank;      // [slot - slot_begin] rank inside its bin
		uint32_t* sort_bin_count; // camera groups | bounce bins | no-pixel bin | shadow bins
		uint32_t* sh_keys;        // [queue index] (NULL: shadow queue stays in append order)
		uint32_t* sh_rank;
		const uint32_t* order;
		const uint32_t* sh_order;
		float sort_min[3], sort_scale; // Morton cells: cubic, (o - sort_min) * sort_scale in [0, 1)
		uint32_t sort_bits;            // cells per axis = 2^sort_bits (path bins)
		uint32_t sort_shadow_bits;     // the same for the shadow-ray bins
		uint32_t sort_dir_major;       // 1: direction bin major, origin cell minor
		uint32_t sort_dir_bits;        // 0: direction octant; n: octahedral map, 2^n x 2^n bins
		uint32_t sort_camera_bins, sort_bounce_bins, sort_shadow_base;
		uint32_t order_reversed;       // k_trace_paths hands the ordered batches out back to front
	};

	// Bin of a ray for the order pass. Regenerated camera rays keep their tile order (one bin per 32 slots: they are
	// coherent as they are); bounce rays are grouped by the Morton cell of their origin, then by direction bin.
	__device__ __forceinline__ uint32_t spread3(uint32_t v)
	{
		v &= 0x3FFu;
		v = (v | (v << 16)) & 0x030000FFu;
		v = (v | (v << 8)) & 0x0300F00Fu;
		v = (v | (v << 4)) & 0x030C30C3u;
		v = (v | (v << 2)) & 0x09249249u;
		return v;
	}
	__device__ __forceinline__ uint32_t morton_cell(const DFrame& f, const float3 o, const uint32_t bits)
	{
		const float cells = float(1u << bits), top = cells - 1.0f, sc = f.sort_scale * cells;
		const uint32_t cx = uint32_t(fminf(fmaxf((o.x - f.sort_min[0]) * sc, 0.0f), top));
		const uint32_t cy = uint32_t(fminf(fmaxf((o.y - f.sort_min[1]) * sc, 0.0f), top));
		const uint32_t cz = uint32_t(fminf(fmaxf((o.z - f.sort_min[2]) * sc, 0.0f), top));
		return (spread3(cx) << 2) | (spread3(cy) << 1) | spread3(cz);
	}
	__device__ __forceinline__ uint32_t direction_bin(const DFrame& f, const float3 d)
	{
		if (f.sort_dir_bits == 0u) return (uint32_t(d.x < 0.0f) << 2) | (uint32_t(d.y < 0.0f) << 1) | uint32_t(d.z < 0.0f);
		// octahedral map of the direction onto [0, 1)^2, 2^n x 2^n cells of about equal solid angle
		const float inv = 1.0f / fmaxf(fabsf(d.x) + fabsf(d.y) + fabsf(d.z), 1.0e-30f);
		float u = d.x * inv, v = d.z * inv;
		if (d.y < 0.0f)
		{
			const float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f);
			v = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
			u = uu;
		}
		const float cells = float(1u << f.sort_dir_bits);
		const uint32_t iu = uint32_t(fminf(fmaxf((u * 0.5f + 0.5f) * cells, 0.0f), cells - 1.0f));
		const uint32_t iv = uint32_t(fminf(fmaxf((v * 0.5f + 0.5f) * cells, 0.0f), cells - 1.0f));
		return (iu << f.sort_dir_bits) | iv;
	}
	__device__ __forceinline__ uint32_t ray_sort_bin(const DFrame& f, const uint32_t slot, const bool has_pixel,
		const bool camera_ray, const float3 o, const float3 d)
	{
		if (!has_pixel) return f.sort_camera_bins + f.sort_bounce_bins;
		if (camera_ray) return (slot - f.slot_begin) >> 5;
		const uint32_t dir_log2 = f.sort_dir_bits ? 2u * f.sort_dir_bits : 3u;
		if (f.sort_dir_major) return f.sort_camera_bins + ((direction_bin(f, d) << (3u * f.sort_bits)) | morton_cell(f, o, f.sort_bits));
		return f.sort_camera_bins + ((morton_cell(f, o, f.sort_bits) << dir_log2) | direction_bin(f, d));
	}
	// one atomic per ray: its rank inside the bin (lanes of a warp that share a bin share the atomic)
	__device__ __forceinline__ uint32_t bin_rank(uint32_t* bin_count, const uint32_t bin)
	{
		const uint32_t peers = __match_any_sync(__activemask(), bin);
		const uint32_t lane = threadIdx.x & 31u;
		const uint32_t leader = __ffs(peers) - 1u;
		uint32_t base = 0u;
		if (lane == leader) base = atomicAdd(bin_count + bin, uint32_t(__popc(peers)));
		base = __shfl_sync(peers, base, leader);
		return base + __popc(peers & ((1u << lane) - 1u));
	}

	// slot -> pixel: 256-slot chunks cover 16x16 pixels (8 tiles of 8x4, 2 across x 4 down); a warp works through one
	// chunk at a time, so all its rays start in one small screen region
	__device__ __forceinline__ bool slot_to_pixel(const DFrame& f, uint32_t slot, uint32_t& x, uint32_t& y)
	{
		const uint32_t chunk = slot >> 8, tile = (slot >> 5) & 7u, within = slot & 31u;
		// chunk / tiles_x by a multiplication (tiles_x_magic = ceil(2^40 / tiles_x): exact while chunk * tiles_x < 2^40):
		// ncu showed the emulated integer divisions of this function among the top stall sites of k_shade
		const uint32_t cy = uint32_t((uint64_t(chunk) * f.tiles_x_magic) >> 40), cx = chunk - cy * f.tiles_x;
		x = cx * 16u + (tile & 1u) * 8u + (within & 7u);
		y = cy * 16u + (tile >> 1) * 4u + (within >> 3);
		return x < f.cam.width && y >= f.row_begin && y < f.row_end && (f.il_count == 1u || cy % f.il_count == f.il_index);
	}

	__device__ __forceinline__ Stack make_stack(uint2* smem_base)
	{
		Stack st;
		st.set_smem(smem_base + threadIdx.x);
		st.sp = 0;
		return st;
	}

	// ---------------------------------------------------------------- scene repacking (rzb_set_scene)
	struct MeshEntry
	{
		uint32_t node_offset, node_count, tri_offset, base; // base = global index of the mesh root, kNoIndex if empty
	};
	// C-ABI triangle (112 B) -> hot (intersection) + cold (shading) records. The edge vectors are formed exactly as
	// Triangle::closestIntersection forms them first (single fp32 subtractions, cuda_render_parts.cuh:1026-1027).
	__global__ void k_pack_triangles(const rzb_triangle* __restrict__ raw, uint32_t n, float4* __restrict__ hot,
		float4* __restrict__ cold)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		const float* t = reinterpret_cast<const float*>(raw + i); // v[9] n[9] face[3] uv[6] slot
		const float v0x = t[0], v0y = t[1], v0z = t[2];
		const float e1x = fsub(t[3], v0x), e1y = fsub(t[4], v0y), e1z = fsub(t[5], v0z);
		const float e2x = fsub(t[6], v0x), e2y = fsub(t[7], v0y), e2z = fsub(t[8], v0z);
		const uint32_t slot = __float_as_uint(t[27]) & 0x3Fu;
		hot[3 * size_t(i)] = make_float4(v0x, v0y, v0z, e1x);
		hot[3 * size_t(i) + 1] = make_float4(e1y, e1z, e2x, e2y);
		hot[3 * size_t(i) + 2] = make_float4(e2z, __uint_as_float(slot), 0.0f, 0.0f);
		cold[5 * size_t(i)] = make_float4(t[9], t[10], t[11], t[21]);
		cold[5 * size_t(i) + 1] = make_float4(t[12], t[13], t[14], t[22]);
		cold[5 * size_t(i) + 2] = make_float4(t[15], t[16], t[17], t[23]);
		cold[5 * size_t(i) + 3] = make_float4(t[18], t[19], t[20], t[24]);
		cold[5 * size_t(i) + 4] = make_float4(t[25], t[26], 0.0f, 0.0f);
	}
	// per-mesh node arrays -> one global array (children / triangle ranges rebased)
	__global__ void k_pack_mesh_nodes(const rzb_node* __restrict__ raw, uint32_t n, const MeshEntry* __restrict__ meshes,
		uint32_t n_meshes, float4* __restrict__ nodes)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		// the mesh whose node range holds i (ranges ascending and disjoint: binary search on node_offset)
		uint32_t lo = 0, hi = n_meshes;
		while (hi - lo > 1u)
		{
			const uint32_t mid = (lo + hi) >> 1;
			if (meshes[mid].node_offset <= i) lo = mid; else hi = mid;
		}
		if (n_meshes == 0u) return;
		const MeshEntry m = meshes[lo];
		if (i < m.node_offset || i >= m.node_offset + m.node_count) return; // a node no mesh owns
		rzb_node nd = raw[i];
		const uint32_t count = nd.type_count & 0x3FFFFFFFu;
		nd.begin += count != 0u ? m.tri_offset : m.base;
		const size_t g = size_t(m.base) + (i - m.node_offset);
		nodes[2 * g] = make_float4(nd.bb_min[0], nd.bb_min[1], nd.bb_min[2], nd.bb_max[0]);
		nodes[2 * g + 1] = make_float4(nd.bb_max[1], nd.bb_max[2], __uint_as_float(nd.begin), __uint_as_float(nd.type_count));
	}
	__global__ void k_iota(uint32_t* __restrict__ out, uint32_t n)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i < n) out[i] = i;
	}
	__global__ void k_iota_from(uint32_t* __restrict__ out, uint32_t n, uint32_t first)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i < n) out[i] = first + i;
	}

	// ---------------------------------------------------------------- k_reset
	__global__ void __launch_bounds__(128) k_reset(DFrame f, uint32_t world_material)
	{
		const uint32_t slot = f.slot_begin + blockIdx.x * blockDim.x + threadIdx.x;
		uint32_t x = 0, y = 0;
		const bool valid = slot < f.slot_end && slot_to_pixel(f, slot, x, y);
		if (valid)
		{
			V3 o, d;
			camera_simple_ray(f.cam, x, y, o, d);
			f.st_o[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(world_material << kMediumShift));
			f.st_d[slot] = make_float4(d.x, d.y, d.z, 1.0f);
			f.st_c[slot] = make_float2(1.0f, 1.0f);
			const size_t p = size_t(y) * f.cam.width + x;
			f.accum[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			f.depth[p] = 0.0f;
		}
	}

	// ---------------------------------------------------------------- work distribution
	// Persistent warps pull batches of 32 consecutive work items with one atomic. Whole-warp batches of neighbouring
	// slots beat every finer-grained scheme that was measured (see rzb_traverse.cuh header).
	__device__ __forceinline__ uint32_t warp_batch(uint32_t* counter)
	{
		uint32_t base = 0u;
		if ((threadIdx.x & 31u) == 0u) base = atomicAdd(counter, 32u);
		return __shfl_sync(0xFFFFFFFFu, base, 0);
	}

	// Streaming accesses (path state, hit records, accumulator, consumed shadow-queue entries) are read or written once per
	// pass; marked evict-first (ld.global.cs / st.global.cs) they leave the L2 to the nodes and triangles: measured +1 % on
	// the 1M-triangle scene, nothing on the materials scene.
	template <class T> __device__ __forceinline__ T ld_s(const T* p) { return __ldcs(p); }
	template <class T> __device__ __forceinline__ void st_s(T* p, const T v) { __stcs(p, v); }

	// ---------------------------------------------------------------- k_trace_paths
	__device__ __forceinline__ void flush_counters(const TraceCounters& cnt, unsigned long long* dst)
	{
		// one atomic per warp and counter
		unsigned long long v[4] = {cnt.top_nodes, cnt.instances, cnt.mesh_nodes, cnt.triangles};
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			unsigned long long x = v[k];
			for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
			if ((threadIdx.x & 31u) == 0u && x) atomicAdd(dst + k, x);
		}
	}

	// RZB_FLAG_COUNT_WORK: lane work and 32 x batch maximum of one whole-warp batch (rzb_work_counters)
	__device__ __forceinline__ void batch_utilisation(const uint32_t work, unsigned long long* out)
	{
		const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, work), mx = __reduce_max_sync(0xFFFFFFFFu, work);
		if ((threadIdx.x & 31u) == 0u)
		{
			atomicAdd(out, (unsigned long long)sum);
			atomicAdd(out + 1, 32ull * mx);
		}
	}

#ifndef RZB_SHADOW_BLOCKS
#define RZB_SHADOW_BLOCKS 8
#endif
#ifndef RZB_TRACE_BLOCKS
#define RZB_TRACE_BLOCKS 7
#endif
	// MODE: 0 = free-running lanes (default), 1 = warp-synchronised rounds (rzb_traverse.cuh: trace_ray)
	template <bool STATS, bool FAST, int MODE = 0, bool WIDE = false>
	__global__ void __launch_bounds__(kTraceBlock, (FAST && !WIDE) ? 8 : RZB_TRACE_BLOCKS) k_trace_paths(DScene sc, DFrame f)
	{
		__shared__ uint2 smem_stack[kSmemStack * kTraceBlock];
		__shared__ ParkedRay smem_park[kTraceBlock];
		Stack st = make_stack(smem_stack);
		ParkedRay& park = smem_park[threadIdx.x];
		TraceCounters cnt{0u, 0u, 0u, 0u};
		for (;;)
		{
			uint32_t base = f.slot_begin + warp_batch(&f.counters[0]);
			if (base >= f.slot_end) break;
			// with ray ordering the bounce rays (the long walks) sit behind the camera rays in the order: hand the batches
			// out from the back, so that the kernel's tail consists of short walks
			if (f.order != nullptr && f.order_reversed) base = f.slot_begin + ((f.slot_end - f.slot_begin - 1u) & ~31u) - (base - f.slot_begin);
			uint32_t slot = base + (threadIdx.x & 31u);
			if (f.order != nullptr && slot < f.slot_end) slot = ld_s(f.order + (slot - f.slot_begin));
			uint32_t x, y;
			const bool active = slot < f.slot_end && slot_to_pixel(f, slot, x, y);
			float4 so = make_float4(0.0f, 0.0f, 0.0f, 0.0f), sd = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
			if (active) { so = ld_s(f.st_o + slot); sd = ld_s(f.st_d + slot); }
			const uint32_t bits = __float_as_uint(so.w);
			const uint32_t depth = bits & 0xFFu, medium = active ? (bits >> kMediumShift) : sc.world_material;
			float near_ = 0.0f, far_ = kFltMax;
			if (depth == 0u) { near_ = f.cam.near_; far_ = f.cam.far_; }
			uint32_t flags = 0u;
			// World::closestIntersection: free flight in the current medium first (cuda_material.cuh:141-159)
			if (active && !(sc.flags & RZB_FLAG_CPU_SEMANTICS))
			{
				const float sigma = sc.materials[medium].scattering;
				if (sigma > 1.0e-4f)
				{
					Rng rng(f.seed, slot, f.pass_index);
					const float dist = (-__logf(rng.next() + 1.0e-4f)) / sigma;
					if (dist < far_) { far_ = dist; flags |= kHitScatterBit; }
				}
			}
			RayResult r;
			trace_ray<false, STATS, MODE, FAST, WIDE>(sc, active, v3(so.x, so.y, so.z), v3(sd.x, sd.y, sd.z), near_, far_, st, park, cnt, r);
			if (STATS) batch_utilisation(active ? r.steps + r.tris : 0u, f.work + 10);
			if (!active) continue;
			uint32_t tri_bits = flags | (r.external ? kHitExternalBit : 0u);
			tri_bits |= (r.tri == kNoIndex) ? kHitTriMask : (r.tri & kHitTriMask);
			st_s(f.hit_a + slot, make_float4(r.t, r.b1, r.b2, __uint_as_float(tri_bits)));
			st_s(f.hit_inst + slot, r.inst);
		}
		if (STATS) flush_counters(cnt, f.work);
	}

	// ---------------------------------------------------------------- multi-ray-per-lane closest hit (rzb_traverse_mr.cuh)
	// The warp loop shared by k_trace_paths_mr and k_trace_rays_mr. `Source` hands out work items and takes results:
	//   uint32_t total() const; uint32_t* counter() const;
	//   bool load(idx, handle, o, d, near, far, user)   false: the item has no ray (a slot without a pixel)
	//   void store(handle, user, result)
	template <int K, int STEPS, bool STATS, bool FAST, class Source>
	__device__ __forceinline__ void mr_run(const DScene& sc, const Source& src, float4* smem, unsigned long long* work_out,
		unsigned long long* occupancy_out)
	{
		MrHot<K> hot{smem + threadIdx.x};
		MrCold<K> cold;
		MrLane lane;
		lane.init(K);
		TraceCounters cnt{0u, 0u, 0u, 0u};
		const uint32_t lane_id = threadIdx.x & 31u;
		const uint32_t n = src.total();
		bool work_left = true;
		uint32_t occ_active = 0u, occ_rounds = 0u;
		for (;;)
		{
			const uint32_t phase = mr_vote(lane, work_left);
			if (phase == kMrDead) break;
			int k = -1;
			if (phase == kMrNode)
			{
				k = lane.pick(kMrNode);
				if (k >= 0) mr_node<K, FAST, STATS, STEPS>(sc, hot, cold, lane, k, cnt);
			}
			else if (phase == kMrLeaf)
			{
				k = lane.pick(kMrLeaf);
				if (k >= 0) mr_leaf<K, FAST, STATS>(sc, hot, cold, lane, k, cnt);
			}
			else if (phase == kMrHeavy)
			{
				k = lane.pick(kMrHeavy);
				if (k >= 0) mr_heavy<K, FAST, STATS>(sc, hot, cold, lane, k, cnt);
			}
			else
			{
				// F: write the finished ray's record, then pull the next work item (one atomic per warp and round)
				k = lane.pick(kMrDone);
				if (k >= 0)
				{
					RayResult r;
					mr_result<K>(hot, cold, k, r);
					src.store(cold.handle[k], cold.user[k], r);
					lane.move(k, kMrDone, kMrEmpty);
				}
				else if (work_left) k = lane.pick(kMrEmpty);
				const bool want = k >= 0 && work_left;
				const uint32_t wanting = __ballot_sync(0xFFFFFFFFu, want);
				uint32_t base = 0u;
				if (wanting != 0u)
				{
					const uint32_t leader = __ffs(wanting) - 1u;
					if (lane_id == leader) base = atomicAdd(src.counter(), uint32_t(__popc(wanting)));
					base = __shfl_sync(0xFFFFFFFFu, base, leader);
					if (base + uint32_t(__popc(wanting)) >= n) work_left = false;
				}
				if (want)
				{
					const uint32_t idx = base + __popc(wanting & ((1u << lane_id) - 1u));
					if (idx < n)
					{
						V3 o, d;
						float near_, far_;
						uint32_t handle, user;
						if (src.load(idx, handle, o, d, near_, far_, user))
						{
							cold.handle[k] = handle; cold.user[k] = user;
							mr_begin<K, FAST, STATS>(sc, hot, cold, lane, k, o, d, near_, far_, cnt);
						}
					}
					else lane.move(k, kMrEmpty, kMrDead);
				}
			}
			lane.rr = (lane.rr + 1u) & 3u;
			if (STATS)
			{
				occ_active += __popc(__ballot_sync(0xFFFFFFFFu, k >= 0));
				occ_rounds += 32u;
			}
		}
		if (STATS)
		{
			flush_counters(cnt, work_out);
			if (lane_id == 0u && occupancy_out)
			{
				atomicAdd(occupancy_out, (unsigned long long)occ_active);
				atomicAdd(occupancy_out + 1, (unsigned long long)occ_rounds);
			}
		}
	}

	struct PathSource
	{
		const DScene* sc;
		const DFrame* f;
		__device__ __forceinline__ uint32_t total() const { return f->slot_end - f->slot_begin; }
		__device__ __forceinline__ uint32_t* counter() const { return &f->counters[0]; }
		__device__ __forceinline__ bool load(const uint32_t idx, uint32_t& slot, V3& o, V3& d, float& near_, float& far_, uint32_t& flags) const
		{
			slot = f->order != nullptr ? f->order[idx] : f->slot_begin + idx;
			uint32_t x, y;
			if (!slot_to_pixel(*f, slot, x, y)) return false;
			const float4 so = f->st_o[slot], sd = f->st_d[slot];
			const uint32_t bits = __float_as_uint(so.w);
			const uint32_t depth = bits & 0xFFu, medium = bits >> kMediumShift;
			near_ = 0.0f; far_ = kFltMax;
			if (depth == 0u) { near_ = f->cam.near_; far_ = f->cam.far_; }
			flags = 0u;
			// World::closestIntersection: free flight in the current medium first (cuda_material.cuh:141-159)
			if (!(sc->flags & RZB_FLAG_CPU_SEMANTICS))
			{
				const float sigma = sc->materials[medium].scattering;
				if (sigma > 1.0e-4f)
				{
					Rng rng(f->seed, slot, f->pass_index);
					const float dist = (-__logf(rng.next() + 1.0e-4f)) / sigma;
					if (dist < far_) { far_ = dist; flags |= kHitScatterBit; }
				}
			}
			o = v3(so.x, so.y, so.z); d = v3(sd.x, sd.y, sd.z);
			return true;
		}
		__device__ __forceinline__ void store(const uint32_t slot, const uint32_t flags, const RayResult& r) const
		{
			uint32_t tri_bits = flags | (r.external ? kHitExternalBit : 0u);
			tri_bits |= (r.tri == kNoIndex) ? kHitTriMask : (r.tri & kHitTriMask);
			f->hit_a[slot] = make_float4(r.t, r.b1, r.b2, __uint_as_float(tri_bits));
			f->hit_inst[slot] = r.inst;
		}
	};

	template <int K, int STEPS, bool STATS, bool FAST>
	__global__ void __launch_bounds__(kMrBlock, 5) k_trace_paths_mr(DScene sc, DFrame f)
	{
		extern __shared__ float4 mr_smem[]; // [kMrFields][K][kMrBlock] (+ padding that only limits blocks per SM)
		const PathSource src{&sc, &f};
		mr_run<K, STEPS, STATS, FAST>(sc, src, mr_smem, f.work, f.work + 10);
	}

	// ---------------------------------------------------------------- shadow queue append (warp-ballot compaction)
	__device__ __forceinline__ void shadow_push(const DFrame& f, bool want, float3 o, float3 d, float dist,
		uint32_t pixel, float3 contrib, uint32_t light_class)
	{
		const uint32_t active = __activemask();
		const uint32_t ballot = __ballot_sync(active, want);
		if (ballot == 0u) return;
		const uint32_t lane = threadIdx.x & 31u;
		const uint32_t leader = __ffs(ballot) - 1u;
		uint32_t base = 0;
		if (lane == leader) base = atomicAdd(&f.counters[1], uint32_t(__popc(ballot)));
		base = __shfl_sync(active, base, leader);
		if (!want) return;
		const uint32_t idx = base + __popc(ballot & ((1u << lane) - 1u));
		if (idx >= f.shadow_capacity) return;
		f.sh_o[idx] = make_float4(o.x, o.y, o.z, dist);
		f.sh_d[idx] = make_float4(d.x, d.y, d.z, __uint_as_float(pixel));
		f.sh_c[idx] = make_float4(contrib.x, contrib.y, contrib.z, 0.0f);
		if (f.sh_keys != nullptr)
		{
			// shadow rays of one light leave a surface cell in (nearly) one direction: cell-major, light class minor
			const uint32_t bin = f.sort_shadow_base + ((morton_cell(f, o, f.sort_shadow_bits) << 2) | (light_class & 3u));
			f.sh_keys[idx] = bin;
			f.sh_rank[idx] = bin_rank(f.sort_bin_count, bin);
		}
	}

	__device__ __forceinline__ bool all_finite(const float3 a, const float3 b, const float3 c)
	{
		// a sum of finite terms can overflow, so test the largest magnitude
		const float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(b.x))),
			fmaxf(fmaxf(fmaxf(fabsf(b.y), fabsf(b.z)), fmaxf(fabsf(c.x), fabsf(c.y))), fabsf(c.z)));
		// fmaxf drops NaN operands: catch them through the sum
		return m < 3.0e38f && (a.x + a.y + a.z + b.x + b.y + b.z + c.x + c.y + c.z) == (a.x + a.y + a.z + b.x + b.y + b.z + c.x + c.y + c.z);
	}

	// ---------------------------------------------------------------- k_shade
#ifndef RZB_SHADE_BLOCKS
#define RZB_SHADE_BLOCKS 6
#endif
	constexpr int kShadeBlocks = RZB_SHADE_BLOCKS; // resident blocks per SM the register allocation aims at
	__global__ void __launch_bounds__(128, kShadeBlocks) k_shade(DScene sc, DFrame f)
	{
		const uint32_t slot = f.slot_begin + blockIdx.x * blockDim.x + threadIdx.x;
		uint32_t x = 0, y = 0;
		const bool valid = slot < f.slot_end && slot_to_pixel(f, slot, x, y);
		// threads without a pixel still take part in the warp-level queue appends
		float4 so = make_float4(0, 0, 0, 0), sd = make_float4(0, 0, 1, 0), ha = make_float4(0, 0, 0, 0);
		float2 scol = make_float2(0, 0);
		uint32_t hinst = kNoIndex;
		if (valid)
		{
			// the pixel's accumulator is read-modified-written at the very end: start bringing it in now
			asm volatile("prefetch.global.L2 [%0];" ::"l"(f.accum + (size_t(y) * f.cam.width + x)));
			so = ld_s(f.st_o + slot); sd = ld_s(f.st_d + slot); scol = ld_s(f.st_c + slot);
			ha = ld_s(f.hit_a + slot); hinst = ld_s(f.hit_inst + slot);
		}
		const uint32_t bits = __float_as_uint(so.w);
		uint32_t depth = bits & 0xFFu;
		uint32_t medium = valid ? (bits >> kMediumShift) : sc.world_material;
		float3 ro = f3(so.x, so.y, so.z);
		float3 rd = f3(sd.x, sd.y, sd.z);
		float3 thr = f3(sd.w, scol.x, scol.y);
		const float far_ = ha.x;
		const uint32_t tri_bits = __float_as_uint(ha.w);
		const bool cpu_sem = (sc.flags & RZB_FLAG_CPU_SEMANTICS) != 0u;

		Rng rng(f.seed, slot, f.pass_index);
		rng.dim = 1u; // dimension 1 belongs to the free-flight draw in k_trace_paths

		// ---- traceRay (cuda_render_kernel.cu:146-237)
		Surface s;
		s.surface_material = s.behind_material = sc.world_material;
		s.u = s.v = 0.0f;
		s.normal = s.mapped_normal = f3(0.0f, 0.0f, 1.0f);
		s.metalness = s.roughness = s.emission = 0.0f;
		s.fresnel = 1.0f; s.reflectance = 0.0f; s.tint_factor = 0.0f; s.refr_x = s.refr_y = 0.0f;
		bool any_hit = false;
		if (tri_bits & kHitScatterBit)
		{
			s.surface_material = s.behind_material = medium;
			s.normal = s.mapped_normal = rd;
			any_hit = true;
		}
		if (valid && hinst != kNoIndex)
		{
			analyze_intersection(sc, hinst, tri_bits & kHitTriMask, ha.y, ha.z, (tri_bits & kHitExternalBit) != 0u, s);
			any_hit = true;
		}
		const rzb_material& smat = sc.materials[s.surface_material];
		if (!(valid && hinst != kNoIndex))
		{
			// sky sphere coordinates (cuda_world.cuh:121-126); only read by map fetches
			const bool mapped = (smat.texture & smat.emission_map & smat.metalness_map & smat.roughness_map) != kNoIndex;
			if (mapped)
			{
				s.u = -(0.5f + atan2f(rd.z, rd.x) * (1.0f / 6.2831853f));
				s.v = 0.5f + asinf(fminf(fmaxf(rd.y, -1.0f), 1.0f)) * (1.0f / 3.14159265f);
			}
		}
		{
			const float4 oc = material_opacity_color(sc, smat, s.u, s.v);
			s.color = f3(oc.x, oc.y, oc.z);
			s.color_alpha = oc.w;
			s.emission = material_emission(sc, smat, s.u, s.v);
		}
		if (!cpu_sem)
		{
			// Beer-Lambert through the current medium (cuda_render_kernel.cu:174-176)
			const rzb_material& mm = sc.materials[medium];
			const float a = 1.0f - mm.color[3];
			const float k = __powf(a, far_);
			thr = thr * f3(mm.color[0], mm.color[1], mm.color[2]) * k;
		}
		float3 final_color = f3(0.0f, 0.0f, 0.0f);
		if (s.emission > 0.0f) final_color = thr * s.color * s.emission;

		bool path_continues = false;
		float3 next_o = ro, next_d = rd;
		if (any_hit && valid)
		{
			++depth;
			s.metalness = material_metalness(sc, smat, s.u, s.v);
			s.roughness = material_roughness(sc, smat, s.u, s.v);
			s.fresnel = fresnel_specular_ratio(s.mapped_normal, rd, sc.materials[medium].ior,
				sc.materials[s.behind_material].ior, s.refr_x, s.refr_y);
			s.reflectance = lerpf(s.fresnel, 1.0f, s.metalness);

			uint32_t next_medium = medium;
			next_d = sample_direction(sc, s, rd, next_medium, rng);
			next_o = ro + rd * far_ + s.normal * (0.0001f * far_);
			medium = next_medium;

			// ---- next event estimation with MIS (cuda_render_kernel.cu:239-355)
			const bool do_direct = sc.direct_light_count != 0u && f.direct_samples != 0u;
			const bool do_spot = sc.spot_light_count != 0u && f.spot_samples != 0u;
			if (do_direct || do_spot)
			{
				const float scattering = smat.scattering;
				const float vS_pdf = brdf(s, scattering, rd, next_d);
				const float3 brdf_color = lerp3(s.color, f3(1.0f, 1.0f, 1.0f), s.reflectance);
				const float3 carry = thr * lerp3(f3(1.0f, 1.0f, 1.0f), s.color, s.metalness);
				const uint32_t pixel = y * f.cam.width + x;
				if (do_direct)
				{
					const float inv_pdf = f.inv_pdf_direct; // float(direct_light_count) / float(direct_samples), divided on the host
					for (uint32_t i = 0; i < f.direct_samples; ++i)
					{
						const uint32_t li = min(uint32_t(rng.next() * float(sc.direct_light_count)), sc.direct_light_count - 1u);
						const rzb_direct_light& L = sc.direct_lights[li];
						const float3 ldir = f3(-L.direction[0], -L.direction[1], -L.direction[2]);
						const float cos_size = __cosf(L.angular_size);
						float Se = 0.0f;
						float3 vPL;
						if (dot(next_d, ldir) > cos_size) { Se = L.emission; vPL = next_d; }
						else
						{
							const float r1 = rng.next(), r2 = rng.next();
							vPL = sample_sphere(r1, r2 * 0.5f * (1.0f - cos_size), ldir);
						}
						const float3 vPLn = normalize(vPL);
						const float b = brdf(s, scattering, rd, vPLn);
						const float solid_angle = 6.2831853f * (1.0f - cos_size);
						const float L_pdf = 1.0f / solid_angle;
						const float vSw = vS_pdf / (vS_pdf + L_pdf);
						const float Lw = 1.0f - vSw;
						const float Le = L.emission * solid_angle * b;
						const float radiance = Le * Lw + Se * vSw;
						const float3 c = f3(L.color[0], L.color[1], L.color[2]) * brdf_color * (radiance * inv_pdf) * carry;
						const bool want = radiance >= 1.0e-4f && all_finite(c, next_o, vPLn);
						shadow_push(f, want, next_o, vPLn, kFltMax, pixel, c, li & 1u);
					}
				}
				if (do_spot)
				{
					const float inv_pdf = f.inv_pdf_spot;
					const float med_scattering = sc.materials[medium].scattering;
					for (uint32_t i = 0; i < f.spot_samples; ++i)
					{
						const uint32_t li = min(uint32_t(rng.next() * float(sc.spot_light_count)), sc.spot_light_count - 1u);
						const rzb_spot_light& L = sc.spot_lights[li];
						const float3 lpos = f3(L.position[0], L.position[1], L.position[2]);
						// SpotLight::sampleDirection (cuda_spot_light.cuh:56-74)
						const float3 vD = normalize(next_d);
						float3 vPL = lpos - next_o;
						float dPL = length(vPL);
						const float vOP_dot_vD = dot(vPL, vD);
						const float dPQ = sqrtf(dPL * dPL - vOP_dot_vD * vOP_dot_vD);
						float Se = 0.0f;
						if (dPQ < L.size && vOP_dot_vD > 0.0f)
						{
							Se = L.emission;
							const float dOQ = sqrtf(dPL * dPL - dPQ * dPQ);
							vPL = next_d * fmaxf(dOQ, 1.0e-4f);
						}
						else
						{
							vPL = sample_disk(vPL * (1.0f / dPL), L.size, rng) + lpos - next_o;
						}
						dPL = length(vPL);
						const float3 vPLn = vPL * (1.0f / dPL);
						const float b = brdf(s, scattering, rd, vPLn);
						const float A = L.size * L.size * 3.14159265f;
						const float d1 = dPL + 1.0f;
						const float solid_angle = A / (d1 * d1);
						const float sctr = __expf(-dPL * med_scattering);
						const float beam = float(__cosf(L.beam_angle) <
							similarity(-vPL, f3(L.direction[0], L.direction[1], L.direction[2])));
						const float L_pdf = 1.0f / solid_angle;
						const float vSw = vS_pdf / (vS_pdf + L_pdf);
						const float Lw = 1.0f - vSw;
						const float Le = L.emission * solid_angle * b;
						const float radiance = (Le * Lw + Se * vSw) * sctr * beam;
						const float3 c = f3(L.color[0], L.color[1], L.color[2]) * brdf_color * (radiance * inv_pdf) * carry;
						const bool want = b >= 1.0e-4f && beam >= 1.0e-4f && radiance >= 1.0e-4f && all_finite(c, next_o, vPLn);
						shadow_push(f, want, next_o, vPLn, dPL, pixel, c, 2u | (li & 1u));
					}
				}
			}
			// ray.color.Blend(ray.color * surface.color, tint_factor)
			thr = thr + (thr * s.color - thr) * s.tint_factor;
			path_continues = depth < f.max_depth;
		}
		// The reference's samplers do not keep directions at unit length (localCoordinate builds an unnormalised basis,
		// the "flip above the surface" step uses the normalised dot product: cuda_render_parts.cuh:1254-1301,
		// cuda_material.cuh:238-240); a chain of scattering bounces can square the length every bounce until the ray
		// overflows (seen once per ~4e8 segments on the materials scene; the reference then keeps a NaN pixel for
		// good). Deliberate difference: a non-finite sample is dropped and its path ended.
		bool discarded = false;
		if (!all_finite(final_color, final_color, final_color)) { final_color = f3(0.0f, 0.0f, 0.0f); discarded = true; }
		if (path_continues && !all_finite(next_o, next_d, thr)) { path_continues = false; discarded = true; }
		// ---- epilogue (cuda_render_kernel.cu:98-120)
		if (valid)
		{
			const size_t p = size_t(y) * f.cam.width + x;
			float4 acc = ld_s(f.accum + p);
			acc.x += final_color.x; acc.y += final_color.y; acc.z += final_color.z;
			acc.w += path_continues ? 0.0f : 1.0f;
			st_s(f.accum + p, acc);
			if (f.pass_index == 0u) f.depth[p] = far_;

			if (!path_continues)
			{
				camera_generate_ray(f.cam, x, y, rng, next_o, next_d);
				medium = sc.world_material;
				thr = f3(1.0f, 1.0f, 1.0f);
				depth = 0u;
			}
			if ((sc.flags & RZB_FLAG_COUNT_WORK) && discarded) atomicAdd(f.work + 9, 1ull); // rzb_work_counters::invalid_rays
			st_s(f.st_o + slot, make_float4(next_o.x, next_o.y, next_o.z, __uint_as_float((medium << kMediumShift) | depth)));
			st_s(f.st_d + slot, make_float4(next_d.x, next_d.y, next_d.z, thr.x));
			st_s(f.st_c + slot, make_float2(thr.y, thr.z));
		}
		if (f.sort_keys != nullptr && slot < f.slot_end)
		{
			const uint32_t bin = ray_sort_bin(f, slot, valid, depth == 0u, next_o, next_d);
			f.sort_keys[slot - f.slot_begin] = bin;
			f.sort_rank[slot - f.slot_begin] = bin_rank(f.sort_bin_count, bin);
		}
	}

	// ---------------------------------------------------------------- k_reproject
	// Camera::reproject + spacialReprojection (cuda_camera.cuh:390-426, cuda_postprocess_kernel.cu:5-16). Runs once per
	// restarted frame, between k_trace_paths and k_shade of pass 0 (the slot still holds the first ray and its hit
	// distance): the hit point is projected into the camera of the frame being replaced and, where that frame's depth
	// agrees within 1 %, its accumulator value times temporal_blend is added.
	__global__ void __launch_bounds__(128) k_reproject(DFrame f)
	{
		const uint32_t slot = f.slot_begin + blockIdx.x * blockDim.x + threadIdx.x;
		uint32_t x = 0, y = 0;
		if (slot >= f.slot_end || !slot_to_pixel(f, slot, x, y)) return;
		const float4 so = f.st_o[slot], sd = f.st_d[slot];
		const float far_ = f.hit_a[slot].x;
		const float3 sp = f3(so.x, so.y, so.z) + f3(sd.x, sd.y, sd.z) * far_;
		const float3 rel = sp - f3(f.prev_cam.px, f.prev_cam.py, f.prev_cam.pz);
		const float lx = f.prev_cam.xx * rel.x + f.prev_cam.xy * rel.y + f.prev_cam.xz * rel.z;
		const float ly = f.prev_cam.yx * rel.x + f.prev_cam.yy * rel.y + f.prev_cam.yz * rel.z;
		const float lz = f.prev_cam.zx * rel.x + f.prev_cam.zy * rel.y + f.prev_cam.zz * rel.z;
		if (!(lz > 0.0f)) return; // behind the previous camera
		const float fx = ((lx / lz) / f.prev_cam.tana + 0.5f) * float(f.cam.width);
		const float fy = ((ly / lz) / (-f.prev_cam.tana / f.cam.aspect) + 0.5f) * float(f.cam.height);
		if (!(fx >= 0.0f && fx < float(f.cam.width) && fy >= 0.0f && fy < float(f.cam.height))) return;
		const size_t q = size_t(uint32_t(fy)) * f.cam.width + uint32_t(fx);
		const float point_dist = length(rel);
		if (!(fabsf(point_dist - f.prev_depth[q]) < 0.01f * point_dist)) return;
		const float4 h = f.prev_accum[q];
		const size_t p = size_t(y) * f.cam.width + x;
		float4 acc = f.accum[p];
		acc.x += h.x * f.reproject_blend; acc.y += h.y * f.reproject_blend;
		acc.z += h.z * f.reproject_blend; acc.w += h.w * f.reproject_blend;
		f.accum[p] = acc;
	}

	// ---------------------------------------------------------------- k_scatter_order
	// bin offsets (prefix sum over DFrame::sort_bin_count) + rank inside the bin -> position in the order arrays
	// Every slot sits in one of the bins in front of the shadow bins, so the shadow bins start at offset n_slots.
	__global__ void __launch_bounds__(256) k_scatter_order(DFrame f, const uint32_t* __restrict__ offsets,
		uint32_t* __restrict__ order, uint32_t* __restrict__ sh_order)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		const uint32_t n_slots = f.slot_end - f.slot_begin;
		if (i < n_slots) order[offsets[f.sort_keys[i]] + f.sort_rank[i]] = f.slot_begin + i;
		if (sh_order != nullptr && i < min(f.counters[1], f.shadow_capacity))
			sh_order[offsets[f.sh_keys[i]] - n_slots + f.sh_rank[i]] = i;
	}

	// ---------------------------------------------------------------- k_trace_shadow
	// Any-hit queries over the shadow queue, whole-warp batches of 32 entries, free-running lanes, CONSERVATIVE box
	// test on every tree (64 registers, 8 blocks per SM). The result of a shadow query is the product of the opacities
	// of all triangles the ray crosses -- independent of visiting order and of which boxes were opened, as long as no
	// box the ray enters is skipped; the widened interval only ever opens more boxes than the reference's exact test.
	// (A triangle in a box that the reference's own test rejects by rounding while the ray grazes it within 4 ulp
	// is counted here and not there; none on the golden shadow rays, tests/test_gpu_parity.py.)
	// Measured, ms per pass, reference trees (materials / 1M triangles): exact decisions + synchronised rounds 1.30 /
	// 0.43; conservative + synchronised 1.12 / 0.35; conservative + free-running 0.77 / 0.25 (kept).
	// Dropped: giving finished lanes new rays between synchronised rounds (threshold 1..24 idle lanes) raised the
	// share of busy lane-rounds from 0.44 to 0.70-0.88 but not the speed (1.28..1.41 ms).
	template <bool STATS, bool WIDE = false>
	__global__ void __launch_bounds__(kTraceBlock, WIDE ? 6 : RZB_SHADOW_BLOCKS) k_trace_shadow(DScene sc, DFrame f)
	{
		__shared__ uint2 smem_stack[kSmemStack * kTraceBlock];
		__shared__ ParkedRay smem_park[kTraceBlock];
		Stack st = make_stack(smem_stack);
		ParkedRay& park = smem_park[threadIdx.x];
		const uint32_t n = min(f.counters[1], f.shadow_capacity);
		TraceCounters cnt{0u, 0u, 0u, 0u};
		if (STATS && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(f.work + 8, (unsigned long long)n);
		for (;;)
		{
			const uint32_t base = warp_batch(&f.counters[2]);
			if (base >= n) break;
			uint32_t i = base + (threadIdx.x & 31u);
			const bool active = i < n;
			if (active && f.sh_order != nullptr) i = ld_s(f.sh_order + i);
			float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
			if (active) { o = ld_s(f.sh_o + i); d = ld_s(f.sh_d + i); }
			RayResult r;
			trace_ray<true, STATS, 0, true, WIDE>(sc, active, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), 0.0f, o.w, st, park, cnt, r);
			if (STATS) batch_utilisation(active ? r.steps + r.tris : 0u, f.work + 12);
			const float w = r.mask.w;
			if (!active || w <= 0.0f) continue;
			const float4 c = ld_s(f.sh_c + i);
			float* a = reinterpret_cast<float*>(f.accum + __float_as_uint(d.w));
			atomicAdd(a + 0, c.x * r.mask.x * w);
			atomicAdd(a + 1, c.y * r.mask.y * w);
			atomicAdd(a + 2, c.z * r.mask.z * w);
		}
		if (STATS) flush_counters(cnt, f.work + 4);
	}

	// ---------------------------------------------------------------- k_tonemap
	struct PeerList
	{
		const float4* accum[8];
		uint32_t count;
	};
	// ComputeFinalColor: divide by the sample count, then three separate multiplications, then c/(c+1);
	// kept in this order (no FMA possible: pure mul/div chain) so RGBA8 truncation matches the oracle bit for bit
	__device__ __forceinline__ uchar4 tonemap_pixel(const float4 p, const float aperture_area, const float exposure_time)
	{
		const float a = p.w == 0.0f ? 1.0f : p.w;
		float r = fdiv(p.x, a), g = fdiv(p.y, a), b = fdiv(p.z, a);
		r = fmul(fmul(fmul(r, aperture_area), exposure_time), 1.0e5f);
		g = fmul(fmul(fmul(g, aperture_area), exposure_time), 1.0e5f);
		b = fmul(fmul(fmul(b, aperture_area), exposure_time), 1.0e5f);
		r = fdiv(r, fadd(r, 1.0f)); g = fdiv(g, fadd(g, 1.0f)); b = fdiv(b, fadd(b, 1.0f));
		return make_uchar4((unsigned char)(fmul(r, 255.0f)), (unsigned char)(fmul(g, 255.0f)), (unsigned char)(fmul(b, 255.0f)), 255);
	}
	__global__ void k_tonemap(const float4* __restrict__ accum, PeerList peers, uchar4* __restrict__ rgba,
		uint32_t n_pixels, float aperture_area, float exposure_time)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n_pixels) return;
		float4 p = accum[i];
		for (uint32_t k = 0; k < peers.count; ++k)
		{
			const float4 q = peers.accum[k][i]; // peer-mapped address: the load crosses NVLink
			p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
		}
		rgba[i] = tonemap_pixel(p, aperture_area, exposure_time);
	}
	// ---------------------------------------------------------------- k_resolve_sliced
	// The exchange step of the one-process-per-GPU path as ONE kernel per rank, no NCCL call and no host round trip:
	//   1. barrier in   every rank tells every peer "my passes are done" by storing the call's epoch into the peer's
	//                   exchange header over NVLink; every block waits until all ranks have arrived in ITS OWN header
	//   2. slice        rank r sums pixels [begin, end) of ALL ranks' accumulators (peer loads over NVLink: all-to-all,
	//                   every GPU pulls (N-1)/N of one frame instead of rank 0 pulling N-1 frames), tone-maps them and
	//                   stores the RGBA8 pixels into the root's staging image (peer stores for r != 0)
	//   3. barrier out  the last block of the rank tells every peer "I am done with your accumulator and my pixels are in
	//                   the root's image" and waits for the same from all peers: when the kernel ends, peers may touch
	//                   their accumulators again and the root may copy the image to the host.
	// Flags only ever grow (epoch = number of the call), so nothing has to be reset between calls.
	struct ExchangeHeader
	{
		uint32_t arrive[8];   // [r]: last epoch rank r has finished rendering for
		uint32_t done[8];     // [r]: last epoch rank r has finished reading / writing for
		uint32_t finished_blocks;
		uint32_t timed_out;   // a peer never showed up (spin limit): the frame is incomplete
		uint32_t _pad[46];
	};
	static_assert(sizeof(ExchangeHeader) == 256, "ExchangeHeader");
	struct SlicedArgs
	{
		const float4* accum[8];    // all ranks, this rank's own first-hand pointer at [rank]
		ExchangeHeader* header[8]; // all ranks
		uchar4* root_rgba;         // staging image in the root's exchange buffer
		uint32_t rank, world, epoch;
		uint32_t begin, end;       // this rank's pixel slice
		float aperture_area, exposure_time;
		unsigned long long spin_limit; // clock64 ticks
	};
	__device__ __forceinline__ void store_flag_sys(uint32_t* p, uint32_t v)
	{
		asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
	}
	__device__ __forceinline__ uint32_t load_flag_sys(const uint32_t* p)
	{
		uint32_t v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
		return v;
	}
	__device__ __forceinline__ bool wait_flags(const uint32_t* flags, uint32_t world, uint32_t epoch, unsigned long long limit)
	{
		// threads 0..world-1 each watch one flag of this rank's own header
		bool ok = true;
		if (threadIdx.x < world)
		{
			const long long t0 = clock64();
			while (int32_t(load_flag_sys(flags + threadIdx.x) - epoch) < 0)
			{
				if ((unsigned long long)(clock64() - t0) > limit) { ok = false; break; }
				__nanosleep(100);
			}
		}
		return ok;
	}
	__global__ void __launch_bounds__(256) k_resolve_sliced(SlicedArgs a)
	{
		ExchangeHeader* own = a.header[a.rank];
		if (blockIdx.x == 0 && threadIdx.x < a.world)
		{
			__threadfence_system();
			store_flag_sys(&a.header[threadIdx.x]->arrive[a.rank], a.epoch);
		}
		if (!wait_flags(own->arrive, a.world, a.epoch, a.spin_limit)) own->timed_out = 1u;
		__syncthreads();
		for (uint32_t i = a.begin + blockIdx.x * blockDim.x + threadIdx.x; i < a.end; i += gridDim.x * blockDim.x)
		{
			float4 p = __ldcv(a.accum[a.rank] + i);
			for (uint32_t r = 0; r < a.world; ++r)
			{
				if (r == a.rank) continue;
				const float4 q = __ldcv(a.accum[r] + i); // peer-mapped address: the load crosses NVLink
				p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
			}
			a.root_rgba[i] = tonemap_pixel(p, a.aperture_area, a.exposure_time);
		}
		__threadfence_system();
		__syncthreads();
		__shared__ bool last;
		if (threadIdx.x == 0) last = atomicAdd(&own->finished_blocks, 1u) == gridDim.x - 1u;
		__syncthreads();
		if (!last) return;
		if (threadIdx.x == 0) own->finished_blocks = 0u;
		if (threadIdx.x < a.world)
		{
			__threadfence_system();
			store_flag_sys(&a.header[threadIdx.x]->done[a.rank], a.epoch);
		}
		if (!wait_flags(own->done, a.world, a.epoch, a.spin_limit)) own->timed_out = 1u;
	}

	// sum of the accumulator's alpha channel (completed paths per pixel): the "spp" a host stops at
	__global__ void __launch_bounds__(256) k_sum_alpha(const float4* __restrict__ accum, uint32_t n, double* __restrict__ out)
	{
		double s = 0.0;
		for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += double(accum[i].w);
		for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, off);
		__shared__ double warp_sum[8];
		if ((threadIdx.x & 31u) == 0u) warp_sum[threadIdx.x >> 5] = s;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			double t = 0.0;
			for (int k = 0; k < 8; ++k) t += warp_sum[k];
			atomicAdd(out, t);
		}
	}
	__global__ void k_accum_add(float4* __restrict__ accum, const float4* __restrict__ other, uint32_t n)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		float4 p = accum[i];
		const float4 q = other[i];
		p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
		accum[i] = p;
	}

	// ---------------------------------------------------------------- ray-set entry points
	struct DHit
	{
		float t, b1, b2;
		uint32_t tri_bits;
		uint32_t inst;
		uint32_t _pad[3];
	};
	static_assert(sizeof(DHit) == 32, "DHit");

	template <bool STATS, bool FAST, bool WIDE = false>
	__global__ void __launch_bounds__(kTraceBlock, 6) k_trace_rays(DScene sc, const float4* __restrict__ ray_o_near,
		const float4* __restrict__ ray_d_far, uint32_t n, DHit* __restrict__ hits, uint32_t* counter,
		unsigned long long* stats)
	{
		__shared__ uint2 smem_stack[kSmemStack * kTraceBlock];
		__shared__ ParkedRay smem_park[kTraceBlock];
		Stack st = make_stack(smem_stack);
		ParkedRay& park = smem_park[threadIdx.x];
		TraceCounters cnt{0u, 0u, 0u, 0u};
		for (;;)
		{
			const uint32_t base = warp_batch(counter);
			if (base >= n) break;
			const uint32_t i = base + (threadIdx.x & 31u);
			const bool active = i < n;
			float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
			if (active) { o = __ldg(ray_o_near + i); d = __ldg(ray_d_far + i); }
			RayResult r;
			trace_ray<false, STATS, false, FAST, WIDE>(sc, active, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, d.w, st, park, cnt, r);
			if (!active) continue;
			const uint32_t tri_bits = (r.tri == kNoIndex ? kHitTriMask : (r.tri & kHitTriMask)) | (r.external ? kHitExternalBit : 0u);
			float4* dst = reinterpret_cast<float4*>(hits + i);
			dst[0] = make_float4(r.t, r.b1, r.b2, __uint_as_float(tri_bits));
			dst[1] = make_float4(__uint_as_float(r.inst), __uint_as_float(STATS ? r.steps : 0u), __uint_as_float(STATS ? r.tris : 0u), 0.0f);
		}
		if (STATS) flush_counters(cnt, stats);
	}

	struct RaySetSource
	{
		const float4* ray_o_near;
		const float4* ray_d_far;
		DHit* hits;
		uint32_t* cnt;
		uint32_t n;
		__device__ __forceinline__ uint32_t total() const { return n; }
		__device__ __forceinline__ uint32_t* counter() const { return cnt; }
		__device__ __forceinline__ bool load(const uint32_t idx, uint32_t& handle, V3& o, V3& d, float& near_, float& far_, uint32_t& user) const
		{
			const float4 ro = __ldg(ray_o_near + idx), rd = __ldg(ray_d_far + idx);
			handle = idx; user = 0u;
			o = v3(ro.x, ro.y, ro.z); d = v3(rd.x, rd.y, rd.z);
			near_ = ro.w; far_ = rd.w;
			return true;
		}
		__device__ __forceinline__ void store(const uint32_t idx, const uint32_t, const RayResult& r) const
		{
			const uint32_t tri_bits = (r.tri == kNoIndex ? kHitTriMask : (r.tri & kHitTriMask)) | (r.external ? kHitExternalBit : 0u);
			float4* dst = reinterpret_cast<float4*>(hits + idx);
			dst[0] = make_float4(r.t, r.b1, r.b2, __uint_as_float(tri_bits));
			dst[1] = make_float4(__uint_as_float(r.inst), 0.0f, 0.0f, 0.0f);
		}
	};
	template <int K, int STEPS, bool STATS, bool FAST>
	__global__ void __launch_bounds__(kMrBlock, 5) k_trace_rays_mr(DScene sc, const float4* __restrict__ ray_o_near,
		const float4* __restrict__ ray_d_far, uint32_t n, DHit* __restrict__ hits, uint32_t* counter, unsigned long long* stats)
	{
		extern __shared__ float4 mr_smem[];
		const RaySetSource src{ray_o_near, ray_d_far, hits, counter, n};
		mr_run<K, STEPS, STATS, FAST>(sc, src, mr_smem, stats, nullptr);
	}

	__global__ void k_convert_hits(DScene sc, const DHit* __restrict__ in, rzb_hit* __restrict__ out, uint32_t n)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		const DHit h = in[i];
		rzb_hit r;
		r.t = h.t;
		if (h.inst == kNoIndex)
		{
			r.instance = RZB_NO_INDEX; r.triangle = RZB_NO_INDEX; r.b1 = 0.0f; r.b2 = 0.0f; r.external = 0u;
		}
		else
		{
			const uint32_t tri = h.tri_bits & kHitTriMask;
			r.instance = sc.inst_host_index[h.inst];
			r.triangle = sc.tri_host_index[tri];
			r.b1 = h.b1; r.b2 = h.b2;
			r.external = (h.tri_bits & kHitExternalBit) ? 1u : 0u;
		}
		out[i] = r;
	}

	// the shadow kernel's traversal flavour (conservative boxes, free-running) over a caller's ray set
	template <bool WIDE = false>
	__global__ void __launch_bounds__(kTraceBlock, WIDE ? 6 : 8) k_trace_any_rays(DScene sc, const float4* __restrict__ ray_o_near,
		const float4* __restrict__ ray_d_far, uint32_t n, float4* __restrict__ masks, uint32_t* counter)
	{
		__shared__ uint2 smem_stack[kSmemStack * kTraceBlock];
		__shared__ ParkedRay smem_park[kTraceBlock];
		Stack st = make_stack(smem_stack);
		ParkedRay& park = smem_park[threadIdx.x];
		TraceCounters cnt{0u, 0u, 0u, 0u};
		for (;;)
		{
			const uint32_t base = warp_batch(counter);
			if (base >= n) break;
			const uint32_t i = base + (threadIdx.x & 31u);
			const bool active = i < n;
			float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
			if (active) { o = __ldg(ray_o_near + i); d = __ldg(ray_d_far + i); }
			RayResult r;
			trace_ray<true, false, false, true, WIDE>(sc, active, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, d.w, st, park, cnt, r);
			if (active) masks[i] = r.mask;
		}
	}

	// pick ray (rayCast kernel, cuda_render_kernel.cu:130-144): pixel-centre ray of the pick pixel with the range
	// depth * [0.99, 1.01]; out[0] = host index of the instance, out[1] = material slot of the triangle
	__global__ void __launch_bounds__(32) k_raycast(DScene sc, DCamera cam, const float* __restrict__ depth, uint32_t px,
		uint32_t py, uint32_t* __restrict__ out)
	{
		__shared__ uint2 smem_stack[kSmemStack * kTraceBlock];
		__shared__ ParkedRay smem_park[32];
		Stack st = make_stack(smem_stack);
		TraceCounters cnt{0u, 0u, 0u, 0u};
		const bool active = threadIdx.x == 0;
		V3 o, d;
		camera_simple_ray(cam, px, py, o, d);
		const float z = depth[size_t(py) * cam.width + px];
		RayResult r;
		trace_ray<false, false, true>(sc, active, o, d, z * 0.99f, z * 1.01f, st, smem_park[threadIdx.x], cnt, r);
		if (!active) return;
		out[0] = RZB_NO_INDEX;
		out[1] = RZB_NO_INDEX;
		if (r.inst != kNoIndex)
		{
			out[0] = sc.inst_host_index[r.inst];
			out[1] = __float_as_uint(__ldg(sc.tri_hot + 3 * size_t(r.tri) + 2).y);
		}
	}

	__global__ void k_camera_rays(DCamera cam, float4* __restrict__ ray_o_near, float4* __restrict__ ray_d_far)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= cam.width * cam.height) return;
		V3 o, d;
		camera_simple_ray(cam, i % cam.width, i / cam.width, o, d);
		ray_o_near[i] = make_float4(o.x, o.y, o.z, cam.near_);
		ray_d_far[i] = make_float4(d.x, d.y, d.z, cam.far_);
	}
}
