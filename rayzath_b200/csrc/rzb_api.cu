// C ABI of the B200 render path (include/rzb200.h): context, world mirror, frame loop, read-back.
// Replaces the host half of the reference's CUDA engine:
//   EngineCore::renderWorld      /root/reference/RayZath/cuda_engine_core.cu:32-128   (mirror + copy-back)
//   Renderer::renderFunction     /root/reference/RayZath/cuda_engine_renderer.cu:73-262 (kernel sequence)
//   World/Mesh/Instance/...::reconstruct  (chunked pinned-memory mirroring -> one flattened upload here)
#include "rzb_kernels.cuh"
#include "rzb_wide.hpp"

#include <nvtx3/nvToolsExt.h> // header-only; ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace rzb;

static_assert(sizeof(rzb_node) == 32, "rzb_node");
static_assert(sizeof(rzb_triangle) == 112, "rzb_triangle");
static_assert(sizeof(rzb_mesh) == 16, "rzb_mesh");
static_assert(sizeof(rzb_instance) == 100, "rzb_instance");
static_assert(sizeof(rzb_material) == 64, "rzb_material");
static_assert(sizeof(rzb_map) == 56, "rzb_map");
static_assert(sizeof(rzb_direct_light) == 32, "rzb_direct_light");
static_assert(sizeof(rzb_spot_light) == 48, "rzb_spot_light");
static_assert(sizeof(rzb_camera) == 92, "rzb_camera");
static_assert(sizeof(rzb_config) == 24, "rzb_config");
static_assert(sizeof(rzb_hit) == 24, "rzb_hit");
static_assert(sizeof(rzb_scene) == 240, "rzb_scene");

namespace
{
	thread_local std::string g_create_error;

	struct DeviceBuffer
	{
		void* ptr = nullptr;
		size_t bytes = 0;
	};
}

namespace rzb
{
	size_t scanTempBytes(uint32_t n);
	cudaError_t exclusiveScan(void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, uint32_t n, cudaStream_t stream);
}

struct rzb_ctx
{
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;     // the stream everything runs on (own_stream unless rzb_set_stream gave another)
	cudaStream_t own_stream = nullptr;
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
	cudaEvent_t ev_resolve[2] = {nullptr, nullptr}; // rzb_resolve_async completion, per slot
	uint32_t* h_pick = nullptr;                     // pinned: ray-cast pick of the two slots
	std::vector<cudaEvent_t> ev_stage; // 4 per sampled pass of the last rzb_render call
	uint32_t sampled_passes = 0;
	bool last_had_shadow = false;
	unsigned long long* d_work = nullptr; // RZB_FLAG_COUNT_WORK counters (16 x u64)
	uint64_t counted_segments = 0;
	std::vector<std::pair<std::string, void*>> ipc_open; // opened peer accumulators (handle bytes -> mapped pointer)
	// sliced resolve (rzb_resolve_sliced): exchange buffer = ExchangeHeader + RGBA8 staging image, epoch of the last call
	void* d_exchange = nullptr;
	size_t exchange_pixels = 0;
	uint32_t exchange_epoch = 0;
	cudaEvent_t ev_exchange[2] = {nullptr, nullptr};
	float last_exchange_ms = 0.0f;
	bool exchange_timed = false;
	std::string error;

	// scene mirror: grow-only device buffers, reused across rzb_set_scene calls (no cudaMalloc/cudaFree when the
	// world keeps its size, which is the per-frame case of a host that re-sends a dirty world)
	enum { kBufNodes, kBufTriRaw, kBufHot, kBufCold, kBufMeshNodesRaw, kBufMeshTable, kBufInstances, kBufInstHost,
		kBufTriHost, kBufInstMats, kBufMaterials, kBufMaps, kBufDirect, kBufSpot, kBufNodes4, kBufInstRoot4, kBufCount };
	DeviceBuffer scene_buf[kBufCount];
	std::vector<DeviceBuffer> map_pixel_buf;
	DScene sc{};
	bool has_scene = false, has_camera = false, frame_ready = false;
	rzb_camera cam{};
	rzb_config cfg{1u, 1u, 16u, 0u, 0u};

	std::vector<void*> frame_allocs;
	DFrame frame{};
	uchar4* d_rgba = nullptr;
	uint32_t* d_counters = nullptr;          // [0..2] work counters, [4..5] 64-bit shadow total
	uint32_t shadow_capacity_alloc = 0;
	void* d_shadow[3] = {nullptr, nullptr, nullptr};
	uint32_t row_begin = 0, row_end = 0; // tile split (rzb_set_rows)
	uint32_t il_index = 0, il_count = 1; // interleaved tile split (rzb_set_row_interleave)

	// scratch for ray-set calls
	DeviceBuffer scratch[4];

	uint64_t passes = 0, launches = 0;
	float last_render_ms = 0.0f, last_trace_ms = 0.0f, last_shade_ms = 0.0f, last_shadow_ms = 0.0f;
	int trace_grid = 0, shadow_grid = 0, rays_grid = 0, any_grid = 0, trace_grid_fast = 0, wide_grid = 0;
	// temporal reprojection history (RZB_FLAG_TEMPORAL_REPROJECTION)
	float4* d_prev_accum = nullptr;
	float* d_prev_depth = nullptr;
	size_t prev_pixels = 0;
	bool has_prev = false;
	rzb_camera frame_cam{}, prev_cam{};
	// geometry of the last full rzb_set_scene (RZB_SCENE_KEEP_GEOMETRY)
	bool geom_valid = false;
	std::vector<uint32_t> geom_mesh_base;
	uint32_t geom_top_base = 0, geom_top_capacity = 0, geom_triangle_count = 0, geom_mesh_depth = 0;
	bool geom_own_trees = false;
	bool geom_wide = false;                    // RZB_SCENE_WIDE_TREES: 4-ary mesh trees built by rzb_set_scene
	std::vector<uint32_t> geom_mesh_root4;     // per mesh: reference of its wide root (kWideEmpty = no tree)
	uint32_t geom_wide_depth = 0;
	bool set_carveout = true;      // RZB200_CARVEOUT=0 disables the shared-memory carve-out hint
	bool own_trees = false;        // rzb_scene::flags & RZB_SCENE_OWN_TREES: conservative box tests
	bool debug_sync = false;       // RZB200_DEBUG_SYNC: synchronise after every kernel of rzb_render and name the one that faulted
	// ray ordering between passes (default on; RZB200_SORT=0 switches it off): k_shade bins the next pass's rays (and the
	// shadow rays it queues) by (Morton cell of the origin, direction bin) with one atomic per ray, a prefix sum over the
	// bins and a scatter turn that into the slot order the traversal kernels pull their batches in
	bool sort_enabled = true;
	uint32_t sort_bits = 5;        // RZB200_SORT_BITS: Morton cells per axis = 2^bits
	uint32_t sort_dir_bits = 3;    // RZB200_SORT_DIRBITS: 0 = direction octant (8 bins), n = octahedral map with 2^n x 2^n bins
	bool sort_shadow = true;       // RZB200_SORT_SHADOW=0: leave the shadow queue in append order
	uint32_t sort_shadow_bits = 6; // RZB200_SORT_SHADOW_BITS: Morton cells per axis of the shadow-ray bins
	bool order_reversed = true;    // RZB200_SORT_REVERSE=0: hand the ordered batches out front to back
	bool sort_dir_major = false;   // RZB200_SORT_MAJOR=1: direction bin is the major key, origin cell the minor one
	bool order_valid = false;      // false until the first ordering after a reset
	enum { kSortKeys, kSortRank, kSortOrder, kSortBins, kSortOffsets, kSortTemp, kSortShKeys, kSortShRank, kSortShOrder, kSortBufCount };
	DeviceBuffer sort_buf[kSortBufCount];
	float sort_min[3] = {0.0f, 0.0f, 0.0f}, sort_extent = 0.0f;
	float last_sort_ms = 0.0f;
	// The shadow kernel of pass p runs on `stream2` beside the closest-hit kernel of pass p + 1 (it only adds to the
	// accumulator, which nothing touches before the next k_shade); the two passes use alternating counter sets. Measured
	// +2.4 % / +2.0 % (1M-triangle / materials scene): one kernel tail per pass is filled. RZB200_OVERLAP=0 or
	// RZB_FLAG_SERIAL_STAGES put it back in stream order (exclusive per-stage times).
	uint32_t stage_stride = 0;     // RZB200_STAGE_STRIDE: 0 = automatic (see rzb_render)
	bool overlap = true;
	cudaStream_t stream2 = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	bool shadow_pending = false;
	// closest-hit kernel flavour: one ray per lane (default) or several rays per lane with phase voting
	// (rzb_traverse_mr.cuh, RZB200_TRACE=mr: measured slower on B200, DESIGN.md section 4 -- kept as the measured
	// alternative; RZB200_MR_RAYS = rays per lane, RZB200_MR_BLOCKS caps its resident blocks per SM)
	bool trace_mr = false;
	int trace_mode = 0;            // RZB200_TRACE_MODE: 0 free-running lanes, 1 warp-synchronised rounds (also RZB200_TRACE_SYNC=1)
	int mr_blocks = 0, mr_k = 2, mr_steps = 2;
	int mr_grid[2] = {0, 0};       // [FAST]
	size_t mr_smem = 0;
};

namespace
{
	// NVTX range over one C-ABI call (the reference has wall-clock stage timers only)
	struct NvtxRange
	{
		explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
		~NvtxRange() { nvtxRangePop(); }
	};

	int fail(rzb_ctx* ctx, int code, const std::string& msg)
	{
		if (ctx) ctx->error = msg;
		else g_create_error = msg;
		return code;
	}
	int cudaFail(rzb_ctx* ctx, cudaError_t e, const char* what)
	{
		return fail(ctx, RZB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
	}
#define RZB_CUDA(ctx, call)                                     \
	do                                                          \
	{                                                           \
		const cudaError_t rzb_e_ = (call);                      \
		if (rzb_e_ != cudaSuccess) return cudaFail(ctx, rzb_e_, #call); \
	} while (0)

	struct DeviceGuard
	{
		int prev = -1;
		explicit DeviceGuard(int dev)
		{
			cudaGetDevice(&prev);
			if (prev != dev) cudaSetDevice(dev);
			else prev = -1;
		}
		~DeviceGuard()
		{
			if (prev >= 0) cudaSetDevice(prev);
		}
	};

	void freeAll(std::vector<void*>& v)
	{
		for (void* p : v) cudaFree(p);
		v.clear();
	}

	int ensureScratch(rzb_ctx* ctx, int slot, size_t bytes)
	{
		DeviceBuffer& b = ctx->scratch[slot];
		if (b.bytes >= bytes) return RZB_OK;
		if (b.ptr) cudaFree(b.ptr);
		b.ptr = nullptr;
		b.bytes = 0;
		RZB_CUDA(ctx, cudaMalloc(&b.ptr, bytes));
		b.bytes = bytes;
		return RZB_OK;
	}

	DCamera makeDeviceCamera(const rzb_camera& c)
	{
		DCamera d{};
		d.width = c.width; d.height = c.height;
		d.px = c.position[0]; d.py = c.position[1]; d.pz = c.position[2];
		d.xx = c.axis_x[0]; d.xy = c.axis_x[1]; d.xz = c.axis_x[2];
		d.yx = c.axis_y[0]; d.yy = c.axis_y[1]; d.yz = c.axis_y[2];
		d.zx = c.axis_z[0]; d.zy = c.axis_z[1]; d.zz = c.axis_z[2];
		d.tana = tanf(c.fov * 0.5f); // host libm, as cpu_engine_kernel.cpp:186
		d.aspect = float(c.width) / float(c.height);
		d.near_ = c.near_far[0]; d.far_ = c.near_far[1];
		d.focal_distance = c.focal_distance;
		d.aperture = c.aperture;
		return d;
	}

	int gridFor(rzb_ctx* ctx, const void* kernel, int block)
	{
		int per_sm = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
		// shared memory and L1 share one array per SM: ask for no more shared memory than the resident blocks use, the
		// rest serves the node / triangle fetches (RZB200_CARVEOUT=0 leaves the driver's default)
		cudaFuncAttributes attr{};
		if (ctx->set_carveout && cudaFuncGetAttributes(&attr, kernel) == cudaSuccess && attr.sharedSizeBytes > 0)
		{
			int max_smem = 0;
			cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device);
			const size_t need = size_t(per_sm) * (attr.sharedSizeBytes + 1024);
			if (max_smem > 0)
			{
				const int pct = int(std::min<size_t>(100, (need * 100 + size_t(max_smem) - 1) / size_t(max_smem)));
				cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
				if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
			}
		}
		return ctx->sm_count * per_sm;
	}

	int prepareFrame(rzb_ctx* ctx)
	{
		// (re)allocate per-camera buffers
		freeAll(ctx->frame_allocs);
		ctx->d_rgba = nullptr;
		DFrame& f = ctx->frame;
		f = DFrame{};
		f.cam = makeDeviceCamera(ctx->cam);
		f.tiles_x = (ctx->cam.width + 15u) / 16u;  // 16x16-pixel chunks of 256 slots
		f.tiles_y = (ctx->cam.height + 15u) / 16u;
		f.tiles_x_magic = ((1ull << 40) + f.tiles_x - 1ull) / f.tiles_x;
		f.n_slots = f.tiles_x * f.tiles_y * 256u;
		ctx->row_begin = 0;
		ctx->row_end = ctx->cam.height;
		const size_t n_pixels = size_t(ctx->cam.width) * ctx->cam.height;
		auto alloc = [&](void** p, size_t bytes) -> int {
			RZB_CUDA(ctx, cudaMalloc(p, bytes));
			ctx->frame_allocs.push_back(*p);
			return RZB_OK;
		};
		int rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.st_o), size_t(f.n_slots) * 16))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.st_d), size_t(f.n_slots) * 16))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.st_c), size_t(f.n_slots) * 8))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.hit_a), size_t(f.n_slots) * 16))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.hit_inst), size_t(f.n_slots) * 4))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.accum), n_pixels * 16))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&f.depth), n_pixels * 4))) return rc;
		if ((rc = alloc(reinterpret_cast<void**>(&ctx->d_rgba), n_pixels * 4))) return rc;
		RZB_CUDA(ctx, cudaMemsetAsync(f.accum, 0, n_pixels * 16, ctx->stream));
		RZB_CUDA(ctx, cudaMemsetAsync(f.depth, 0, n_pixels * 4, ctx->stream));
		ctx->frame_ready = false;
		ctx->passes = 0;       // new buffers: nothing rendered, no reprojection history
		ctx->has_prev = false;
		return RZB_OK;
	}

	// pixels this context traces per pass (row band intersected with its interleaved 16-row chunk rows)
	uint64_t bandPixels(const rzb_ctx* ctx)
	{
		const DFrame& f = ctx->frame;
		uint64_t rows = 0;
		for (uint32_t r = f.row_begin / 16u; r * 16u < f.row_end; ++r)
		{
			if (r % std::max(f.il_count, 1u) != f.il_index) continue;
			const uint32_t y0 = std::max(r * 16u, f.row_begin), y1 = std::min(r * 16u + 16u, f.row_end);
			rows += y1 > y0 ? y1 - y0 : 0u;
		}
		return rows * ctx->cam.width;
	}

	void applyRows(rzb_ctx* ctx)
	{
		DFrame& f = ctx->frame;
		f.row_begin = std::min(ctx->row_begin, ctx->cam.height);
		f.row_end = std::min(std::max(ctx->row_end, f.row_begin), ctx->cam.height);
		f.slot_begin = (f.row_begin / 16u) * f.tiles_x * 256u;
		f.slot_end = f.row_end > f.row_begin ? ((f.row_end + 15u) / 16u) * f.tiles_x * 256u : f.slot_begin;
		f.il_index = ctx->il_index;
		f.il_count = std::max(ctx->il_count, 1u);
	}

	int ensureShadowQueue(rzb_ctx* ctx)
	{
		const uint32_t per_pixel =
			(ctx->sc.direct_light_count ? ctx->cfg.direct_light_samples : 0u) +
			(ctx->sc.spot_light_count ? ctx->cfg.spot_light_samples : 0u);
		const uint32_t want = std::max<uint32_t>(ctx->frame.n_slots * std::max(per_pixel, 1u), 32u);
		if (want > ctx->shadow_capacity_alloc)
		{
			for (void*& p : ctx->d_shadow)
			{
				if (p) cudaFree(p);
				p = nullptr;
			}
			for (void*& p : ctx->d_shadow) RZB_CUDA(ctx, cudaMalloc(&p, size_t(want) * 16));
			ctx->shadow_capacity_alloc = want;
		}
		ctx->frame.sh_o = static_cast<float4*>(ctx->d_shadow[0]);
		ctx->frame.sh_d = static_cast<float4*>(ctx->d_shadow[1]);
		ctx->frame.sh_c = static_cast<float4*>(ctx->d_shadow[2]);
		ctx->frame.shadow_capacity = want;
		return RZB_OK;
	}
}

extern "C" int rzb_abi_version(void) { return int(RZB_ABI_VERSION); }

extern "C" const char* rzb_last_error(const rzb_ctx* ctx)
{
	return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

namespace
{
	typedef void (*MrPathKernel)(DScene, DFrame);
	typedef void (*MrRaysKernel)(DScene, const float4*, const float4*, uint32_t, DHit*, uint32_t*, unsigned long long*);
	template <int K, int STEPS>
	MrPathKernel mrPathKernelKS(const bool stats, const bool fast)
	{
		if (stats) return fast ? &k_trace_paths_mr<K, STEPS, true, true> : &k_trace_paths_mr<K, STEPS, true, false>;
		return fast ? &k_trace_paths_mr<K, STEPS, false, true> : &k_trace_paths_mr<K, STEPS, false, false>;
	}
	const void* mrPathKernel(const int k, const int steps, const bool stats, const bool fast)
	{
		(void)steps; // instantiated: 2 or 4 rays per lane, two pair steps per node round
		const MrPathKernel f = k <= 2 ? mrPathKernelKS<2, 2>(stats, fast) : mrPathKernelKS<4, 2>(stats, fast);
		return reinterpret_cast<const void*>(f);
	}
	template <int K, int STEPS>
	MrRaysKernel mrRaysKernelKS(const bool fast) { return fast ? &k_trace_rays_mr<K, STEPS, false, true> : &k_trace_rays_mr<K, STEPS, false, false>; }
	const void* mrRaysKernel(const int k, const int steps, const bool fast)
	{
		(void)steps;
		const MrRaysKernel f = k <= 2 ? mrRaysKernelKS<2, 2>(fast) : mrRaysKernelKS<4, 2>(fast);
		return reinterpret_cast<const void*>(f);
	}
}

extern "C" int rzb_create(int device, rzb_ctx** out)
{
	if (!out) return fail(nullptr, RZB_ERR_INVALID, "rzb_create: out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) return cudaFail(nullptr, e, "cudaGetDeviceCount (no CUDA device: this path has no CPU fallback)");
	if (device < 0 || device >= count) return fail(nullptr, RZB_ERR_INVALID, "rzb_create: device ordinal out of range");
	rzb_ctx* ctx = new (std::nothrow) rzb_ctx();
	if (!ctx) return fail(nullptr, RZB_ERR_NOMEM, "rzb_create: out of host memory");
	ctx->device = device;
	DeviceGuard guard(device);
	cudaDeviceProp prop{};
	if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaGetDeviceProperties"); }
	ctx->sm_count = prop.multiProcessorCount;
	if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaStreamCreate"); }
	ctx->stream = ctx->own_stream;
	if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaStreamCreate"); }
	cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
	cudaEventCreate(&ctx->ev_begin);
	cudaEventCreate(&ctx->ev_end);
	cudaEventCreateWithFlags(&ctx->ev_resolve[0], cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->ev_resolve[1], cudaEventDisableTiming);
	// pinned: [0..1] / [2..3] ray-cast pick of the two asynchronous-resolve slots, [4] "a peer never arrived" of the sliced resolve
	if (cudaMallocHost(reinterpret_cast<void**>(&ctx->h_pick), 32) != cudaSuccess) ctx->h_pick = nullptr;
	else std::memset(ctx->h_pick, 0, 32);
	if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_work), 128)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaMalloc(work)"); }
	cudaMemsetAsync(ctx->d_work, 0, 128, ctx->stream);
	if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_counters), 256)) != cudaSuccess) { delete ctx; return cudaFail(nullptr, e, "cudaMalloc(counters)"); }
	cudaMemsetAsync(ctx->d_counters, 0, 256, ctx->stream);
	ctx->debug_sync = std::getenv("RZB200_DEBUG_SYNC") != nullptr;
	if (const char* env = std::getenv("RZB200_CARVEOUT")) ctx->set_carveout = std::atoi(env) != 0;
	if (const char* env = std::getenv("RZB200_TRACE")) ctx->trace_mr = std::string(env) == "mr";
	if (const char* env = std::getenv("RZB200_TRACE_SYNC")) ctx->trace_mode = std::atoi(env) != 0 ? 1 : 0;
	if (const char* env = std::getenv("RZB200_MR_BLOCKS")) ctx->mr_blocks = std::atoi(env);
	if (const char* env = std::getenv("RZB200_SORT")) ctx->sort_enabled = std::atoi(env) != 0;
	if (const char* env = std::getenv("RZB200_SORT_BITS")) ctx->sort_bits = uint32_t(std::min(std::max(std::atoi(env), 1), 6));
	if (const char* env = std::getenv("RZB200_SORT_DIRBITS")) ctx->sort_dir_bits = uint32_t(std::min(std::max(std::atoi(env), 0), 4));
	if (const char* env = std::getenv("RZB200_SORT_SHADOW")) ctx->sort_shadow = std::atoi(env) != 0;
	if (const char* env = std::getenv("RZB200_SORT_SHADOW_BITS")) ctx->sort_shadow_bits = uint32_t(std::min(std::max(std::atoi(env), 1), 7));
	if (const char* env = std::getenv("RZB200_SORT_MAJOR")) ctx->sort_dir_major = std::atoi(env) != 0;
	if (const char* env = std::getenv("RZB200_SORT_REVERSE")) ctx->order_reversed = std::atoi(env) != 0;
	// RZB200_OVERLAP=0: a pass's shadow kernel in stream order instead of on a second stream beside the next pass's closest-hit
	// kernel
	if (const char* env = std::getenv("RZB200_OVERLAP")) ctx->overlap = std::atoi(env) != 0;
	if (const char* env = std::getenv("RZB200_STAGE_STRIDE")) ctx->stage_stride = uint32_t(std::max(std::atoi(env), 1));
	if (const char* env = std::getenv("RZB200_TRACE_MODE")) ctx->trace_mode = std::atoi(env) == 1 ? 1 : 0;
	ctx->trace_grid = gridFor(ctx, reinterpret_cast<const void*>(&k_trace_paths<false, false>), kTraceBlock);
	ctx->shadow_grid = gridFor(ctx, reinterpret_cast<const void*>(&k_trace_shadow<false>), kTraceBlock);
	ctx->trace_grid_fast = gridFor(ctx, reinterpret_cast<const void*>(&k_trace_paths<false, true>), kTraceBlock);
	ctx->rays_grid = gridFor(ctx, reinterpret_cast<const void*>(&k_trace_rays<false, false>), kTraceBlock);
	ctx->any_grid = gridFor(ctx, reinterpret_cast<const void*>(&k_trace_any_rays<false>), kTraceBlock);
	ctx->wide_grid = gridFor(ctx, reinterpret_cast<const void*>(&k_trace_paths<false, true, false, true>), kTraceBlock);
	gridFor(ctx, reinterpret_cast<const void*>(&k_trace_rays<false, true, true>), kTraceBlock);
	if (const char* env = std::getenv("RZB200_MR_RAYS")) ctx->mr_k = std::atoi(env) <= 2 ? 2 : 4;
	if (const char* env = std::getenv("RZB200_MR_STEPS")) ctx->mr_steps = std::min(std::max(std::atoi(env), 1), 2);
	{
		// multi-ray kernels: dynamic shared memory = the rays' hot state, padded when fewer resident blocks are asked for
		const size_t need = size_t(kMrFields) * size_t(ctx->mr_k) * kMrBlock * sizeof(float4);
		int max_smem = 0;
		cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
		size_t bytes = need;
		if (ctx->mr_blocks > 0 && max_smem > 0)
			bytes = std::max(need, std::min<size_t>(size_t(max_smem) / size_t(ctx->mr_blocks + 1) + 1024, size_t(max_smem) / size_t(ctx->mr_blocks) - 1024));
		ctx->mr_smem = bytes;
		for (int fast = 0; fast < 2; ++fast)
			for (int stats = 0; stats < 2; ++stats)
			{
				const void* kernels[2] = {mrPathKernel(ctx->mr_k, ctx->mr_steps, stats != 0, fast != 0), mrRaysKernel(ctx->mr_k, ctx->mr_steps, fast != 0)};
				for (const void* kernel : kernels)
				{
					if (!kernel) continue;
					cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
					int per_sm = 0;
					if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMrBlock, bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
					if (ctx->set_carveout && max_smem > 0)
					{
						const int pct = int(std::min<size_t>(100, (size_t(per_sm) * (bytes + 1024) * 100 + size_t(max_smem) - 1) / size_t(max_smem)));
						cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
						if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMrBlock, bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
					}
					if (!stats && kernel == kernels[0]) ctx->mr_grid[fast] = ctx->sm_count * per_sm;
				}
			}
		cudaGetLastError();
	}
	*out = ctx;
	return RZB_OK;
}

extern "C" void rzb_destroy(rzb_ctx* ctx)
{
	if (!ctx) return;
	DeviceGuard guard(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (auto& b : ctx->scene_buf) if (b.ptr) cudaFree(b.ptr);
	for (auto& b : ctx->map_pixel_buf) if (b.ptr) cudaFree(b.ptr);
	freeAll(ctx->frame_allocs);
	for (void* p : ctx->d_shadow) if (p) cudaFree(p);
	for (auto& b : ctx->scratch) if (b.ptr) cudaFree(b.ptr);
	if (ctx->d_counters) cudaFree(ctx->d_counters);
	if (ctx->d_work) cudaFree(ctx->d_work);
	if (ctx->d_prev_accum) cudaFree(ctx->d_prev_accum);
	if (ctx->d_prev_depth) cudaFree(ctx->d_prev_depth);
	for (auto& h : ctx->ipc_open) cudaIpcCloseMemHandle(h.second);
	if (ctx->d_exchange) cudaFree(ctx->d_exchange);
	for (auto& b : ctx->sort_buf) if (b.ptr) cudaFree(b.ptr);
	for (cudaEvent_t ev : ctx->ev_exchange) if (ev) cudaEventDestroy(ev);
	cudaEventDestroy(ctx->ev_begin);
	cudaEventDestroy(ctx->ev_end);
	cudaEventDestroy(ctx->ev_resolve[0]);
	cudaEventDestroy(ctx->ev_resolve[1]);
	if (ctx->h_pick) cudaFreeHost(ctx->h_pick);
	for (auto& ev : ctx->ev_stage) cudaEventDestroy(ev);
	cudaStreamDestroy(ctx->own_stream);
	if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
	if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
	if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	delete ctx;
}

extern "C" int rzb_set_stream(rzb_ctx* ctx, void* cuda_stream, int use_caller_stream)
{
	if (!ctx) return RZB_ERR_INVALID;
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->stream = use_caller_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
	return RZB_OK;
}

namespace
{
	// grow-only buffer: reallocates only when the requested size exceeds the capacity
	int ensureBuf(rzb_ctx* ctx, DeviceBuffer& b, size_t bytes)
	{
		bytes = std::max<size_t>(bytes, 16);
		if (b.bytes >= bytes) return RZB_OK;
		if (b.ptr) cudaFree(b.ptr);
		b.ptr = nullptr;
		b.bytes = 0;
		const size_t cap = bytes + bytes / 8; // a little slack so that small growth does not reallocate
		RZB_CUDA(ctx, cudaMalloc(&b.ptr, cap));
		b.bytes = cap;
		return RZB_OK;
	}
	template <typename T>
	int uploadTo(rzb_ctx* ctx, DeviceBuffer& b, const T* host, size_t count, const T** out)
	{
		const int rc = ensureBuf(ctx, b, count * sizeof(T));
		if (rc) return rc;
		if (count) RZB_CUDA(ctx, cudaMemcpyAsync(b.ptr, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
		if (out) *out = static_cast<const T*>(b.ptr);
		return RZB_OK;
	}
}

namespace
{
	// Walks a caller-supplied tree from its root (node 0) and checks what the device traversal relies on: leaves
	// (count != 0) stay inside [0, leaf_limit); inner nodes (count == 0) have their two children as an adjacent pair at
	// an odd index; every node is reached at most once (no cycles, no shared subtrees -- nodes no one references are
	// tolerated). Returns NULL and the depth of the deepest node, or what is wrong.
	const char* validateTree(const rzb_node* nodes, uint32_t node_count, uint32_t leaf_limit, uint32_t& depth_out)
	{
		depth_out = 0;
		if (node_count == 0) return "no nodes";
		std::vector<uint8_t> seen(node_count, 0);
		std::vector<std::pair<uint32_t, uint32_t>> todo;
		todo.emplace_back(0u, 0u);
		seen[0] = 1;
		while (!todo.empty())
		{
			const uint32_t i = todo.back().first, depth = todo.back().second;
			todo.pop_back();
			depth_out = std::max(depth_out, depth);
			const rzb_node& n = nodes[i];
			const uint32_t count = n.type_count & 0x3FFFFFFFu;
			if (count != 0)
			{
				if (uint64_t(n.begin) + count > leaf_limit) return "leaf range outside the object array";
				continue;
			}
			if (uint64_t(n.begin) + 1 >= node_count || (n.begin & 1u) == 0)
				return "bad child index (children are an adjacent pair at an odd index; a leaf has count != 0, an empty tree has no nodes)";
			if (seen[n.begin] || seen[n.begin + 1]) return "a node is referenced twice (cycle or shared subtree)";
			seen[n.begin] = seen[n.begin + 1] = 1;
			todo.emplace_back(n.begin, depth + 1u);
			todo.emplace_back(n.begin + 1u, depth + 1u);
		}
		return nullptr;
	}
}

extern "C" int rzb_set_scene(rzb_ctx* ctx, const rzb_scene* s)
{
	if (!ctx || !s) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: NULL argument");
	NvtxRange nvtx("rzb_set_scene");
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	struct InFlight
	{
		rzb_ctx* ctx;
		bool armed = false;
		~InFlight() { if (armed) cudaStreamSynchronize(ctx->stream); }
	} in_flight{ctx};
	ctx->has_scene = false;
	ctx->frame_ready = false;

	const bool keep_geometry = (s->flags & RZB_SCENE_KEEP_GEOMETRY) != 0u;
	if (keep_geometry)
	{
		if (!ctx->geom_valid) return fail(ctx, RZB_ERR_STATE, "rzb_set_scene: RZB_SCENE_KEEP_GEOMETRY without a previous full upload");
		if (s->instance_node_count > ctx->geom_top_capacity)
			return fail(ctx, RZB_ERR_STATE, "rzb_set_scene: the instance tree outgrew its reserved space, upload the whole scene");
	}
	const uint32_t mesh_count = keep_geometry ? uint32_t(ctx->geom_mesh_base.size()) : s->mesh_count;
	const uint32_t triangle_count = keep_geometry ? ctx->geom_triangle_count : s->triangle_count;

	// ---- validate (host reads of the caller's arrays; everything heavy happens on the device below)
	if (triangle_count > kHitTriMask - 1u) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: too many triangles");
	if (s->default_material >= s->material_count) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: default material out of range");
	for (uint32_t i = 0; i < s->instance_material_count; ++i)
		if (s->instance_materials[i] >= s->material_count)
			return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: instance material id out of range");
	if (s->instance_count != 0 && s->instance_node_count == 0)
		return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: instances without an instance tree");

	// (materials and maps first: the map pixels are the second largest upload and need only the checks below, so their
	// copies run while the host walks the trees further down; `in_flight` waits for them on every way out)
	// ---- materials (+ world material appended), maps
	std::vector<rzb_material> mats(s->materials, s->materials + s->material_count);
	mats.push_back(s->world_material);
	for (const rzb_material& m : mats)
	{
		const uint32_t ids[5] = {m.texture, m.normal_map, m.metalness_map, m.roughness_map, m.emission_map};
		for (uint32_t id : ids)
			if (id != RZB_NO_INDEX && id >= s->map_count) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: map id out of range");
	}
	if (ctx->map_pixel_buf.size() < s->map_count) ctx->map_pixel_buf.resize(s->map_count);
	std::vector<DMap> maps(s->map_count);
	int rc;
	in_flight.armed = true;
	for (uint32_t i = 0; i < s->map_count; ++i)
	{
		const rzb_map& m = s->maps[i];
		if (!m.pixels || m.width == 0 || m.height == 0 || m.format > RZB_MAP_R32F)
			return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: bad map");
		const size_t texel = m.format == RZB_MAP_R8 ? 1 : 4;
		const uint8_t* d_pixels = nullptr;
		if ((rc = uploadTo(ctx, ctx->map_pixel_buf[i], static_cast<const uint8_t*>(m.pixels), size_t(m.width) * m.height * texel, &d_pixels))) return rc;
		DMap d{};
		d.pixels = d_pixels;
		d.width = m.width; d.height = m.height;
		d.format = m.format; d.filter = m.filter; d.address = m.address;
		d.scale_x = m.scale[0]; d.scale_y = m.scale[1];
		d.rot_sin = sinf(m.rotation); d.rot_cos = cosf(m.rotation);
		d.trans_x = m.translation[0]; d.trans_y = m.translation[1];
		maps[i] = d;
	}

	// ---- node placement: every tree is placed so that its root sits at an odd global index; sibling pairs (odd
	// local index, next even) then start at even global indices = 64-byte aligned
	std::vector<WideNode> wide_nodes; // RZB_SCENE_WIDE_TREES: the collapsed mesh trees of a full upload
	const rzb_node* d_mesh_nodes_raw = nullptr;
	std::vector<MeshEntry> table; // non-empty meshes only, ascending node_offset
	std::vector<uint32_t> mesh_base(mesh_count, kNoIndex);
	uint32_t top_base = 0;
	size_t total_nodes = 0;
	uint32_t top_capacity = 0;
	uint32_t mesh_depth = ctx->geom_mesh_depth;
	if (keep_geometry)
	{
		mesh_base = ctx->geom_mesh_base;
		top_base = ctx->geom_top_base;
		top_capacity = ctx->geom_top_capacity;
	}
	else
	{
		ctx->geom_valid = false; // until this upload has succeeded
		mesh_depth = 0;
		table.reserve(s->mesh_count);
		size_t cursor = 1;
		uint64_t expect_node = 0;
		for (uint32_t m = 0; m < s->mesh_count; ++m)
		{
			const rzb_mesh& mesh = s->meshes[m];
			if (uint64_t(mesh.node_offset) + mesh.node_count > s->mesh_node_count ||
				uint64_t(mesh.tri_offset) + mesh.tri_count > s->triangle_count)
				return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: mesh range outside node/triangle arrays");
			if (mesh.node_count != 0 && mesh.node_offset < expect_node)
				return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: mesh node ranges must be disjoint and ascending");
			if (mesh.node_count != 0) expect_node = uint64_t(mesh.node_offset) + mesh.node_count;
			if (mesh.node_count != 0 && mesh.tri_count != 0) // a mesh without triangles has no tree to walk (its root would be a leaf with count 0)
			{
				if ((cursor & 1u) == 0) ++cursor;
				mesh_base[m] = uint32_t(cursor);
				cursor += mesh.node_count;
				table.push_back(MeshEntry{mesh.node_offset, mesh.node_count, mesh.tri_offset, mesh_base[m]});
			}
		}
		if ((cursor & 1u) == 0) ++cursor;
		top_base = uint32_t(cursor);
		// room for the instance tree to grow under RZB_SCENE_KEEP_GEOMETRY updates
		top_capacity = std::max<uint32_t>(2u * s->instance_node_count, 1024u);
		cursor += top_capacity;
		if (cursor >= (1u << 30)) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: too many nodes");
		total_nodes = cursor + 1;
		// The triangles are the bulk of the upload (112 B each) and need no host-side check beyond the ranges above: start
		// their copy and the device repack NOW, so that the DMA runs while the host walks the trees below. `in_flight`
		// waits for the stream on every way out of this function (the caller may free its arrays after an error return).
		{
			int rc0;
			const rzb_triangle* d_tri_raw = nullptr;
			if ((rc0 = uploadTo(ctx, ctx->scene_buf[rzb_ctx::kBufTriRaw], s->triangles, s->triangle_count, &d_tri_raw))) return rc0;
			in_flight.armed = true;
			if ((rc0 = ensureBuf(ctx, ctx->scene_buf[rzb_ctx::kBufHot], size_t(s->triangle_count) * 48))) return rc0;
			if ((rc0 = ensureBuf(ctx, ctx->scene_buf[rzb_ctx::kBufCold], size_t(s->triangle_count) * 80))) return rc0;
			if (s->triangle_count)
			{
				k_pack_triangles<<<(s->triangle_count + 127) / 128, 128, 0, ctx->stream>>>(d_tri_raw, s->triangle_count,
					static_cast<float4*>(ctx->scene_buf[rzb_ctx::kBufHot].ptr), static_cast<float4*>(ctx->scene_buf[rzb_ctx::kBufCold].ptr));
				ctx->launches += 1;
			}
			if ((rc0 = uploadTo(ctx, ctx->scene_buf[rzb_ctx::kBufMeshNodesRaw], s->mesh_nodes, s->mesh_node_count, &d_mesh_nodes_raw))) return rc0;
			if (s->tri_host_index &&
				(rc0 = uploadTo(ctx, ctx->scene_buf[rzb_ctx::kBufTriHost], s->tri_host_index, s->triangle_count, static_cast<const uint32_t**>(nullptr)))) return rc0;
		}
		// every tree must BE a tree (each node reached once from the root: no cycles, no shared subtrees) and the deepest
		// one bounds the traversal stack (rzb_device.cuh: Stack)
		for (uint32_t m = 0; m < s->mesh_count; ++m)
		{
			const rzb_mesh& mesh = s->meshes[m];
			if (mesh_base[m] == kNoIndex) continue;
			uint32_t depth = 0;
			const char* what = validateTree(s->mesh_nodes + mesh.node_offset, mesh.node_count, mesh.tri_count, depth);
			if (what) return fail(ctx, RZB_ERR_INVALID, std::string("rzb_set_scene: mesh tree: ") + what);
			mesh_depth = std::max(mesh_depth, depth);
		}
		ctx->geom_mesh_depth = mesh_depth;
		// ---- optional wide collapse of the own trees
		ctx->geom_wide = false;
		wide_nodes.clear();
		if ((s->flags & RZB_SCENE_WIDE_TREES) != 0u)
		{
			if ((s->flags & RZB_SCENE_OWN_TREES) == 0u)
				return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: RZB_SCENE_WIDE_TREES needs RZB_SCENE_OWN_TREES (the reference's trees are walked as they are)");
			ctx->geom_mesh_root4.assign(s->mesh_count, kWideEmpty);
			ctx->geom_wide_depth = 0;
			bool ok = true;
			for (uint32_t m = 0; m < s->mesh_count && ok; ++m)
				if (mesh_base[m] != kNoIndex)
					ctx->geom_mesh_root4[m] = collapseWide(s->mesh_nodes + s->meshes[m].node_offset, 0u, s->meshes[m].tri_offset, wide_nodes, 0u,
						ctx->geom_wide_depth, ok);
			if (!ok) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: RZB_SCENE_WIDE_TREES supports at most 2^25 triangles and 15 triangles per leaf");
			ctx->geom_wide = true;
		}
	}
	if (ctx->geom_wide && 3u * ctx->geom_wide_depth + 4u > uint32_t(kSmemStack + kLocalStack))
		return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: wide trees too deep for the traversal stack");
	// instance tree (small): fixed up on the host
	// (a world without instances has no tree to walk: its nodes, if any, are not looked at)
	const uint32_t top_node_count = s->instance_count ? s->instance_node_count : 0u;
	std::vector<rzb_node> top_nodes(top_node_count);
	uint32_t top_depth = 0;
	if (top_node_count)
	{
		const char* what = validateTree(s->instance_nodes, top_node_count, s->instance_count, top_depth);
		if (what) return fail(ctx, RZB_ERR_INVALID, std::string("rzb_set_scene: instance tree: ") + what);
	}
	for (uint32_t i = 0; i < top_node_count; ++i)
	{
		rzb_node n = s->instance_nodes[i];
		if ((n.type_count & 0x3FFFFFFFu) == 0u) n.begin += top_base;
		top_nodes[i] = n;
	}
	// one deferred sibling per level of either tree plus the instance-range entry must fit the traversal stack
	if (top_depth + 1u + mesh_depth + 2u > uint32_t(kSmemStack + kLocalStack))
		return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: trees too deep for the traversal stack (instance tree depth + mesh tree depth must stay below "
			+ std::to_string(kSmemStack + kLocalStack - 2) + ")");

	// ---- instances
	std::vector<DInstance> insts(s->instance_count);
	std::vector<uint32_t> inst_host(s->instance_count);
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
		d.mesh_root = kNoIndex;
		if (h.mesh != RZB_NO_INDEX)
		{
			if (h.mesh >= mesh_count) return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: instance mesh id out of range");
			d.mesh_root = mesh_base[h.mesh];
		}
		if (uint64_t(h.material_offset) + h.material_count > s->instance_material_count ||
			h.material_count > RZB_MAX_MATERIALS_PER_INSTANCE)
			return fail(ctx, RZB_ERR_INVALID, "rzb_set_scene: instance material slice out of range");
		d.mat_offset = h.material_offset;
		d.mat_count = h.material_count;
		insts[i] = d;
		inst_host[i] = h.host_index;
	}

	// ---- uploads straight from the caller's arrays, then device-side repacking into the traversal layout
	DScene sc{};
	DeviceBuffer* B = ctx->scene_buf;
	float4* d_nodes = nullptr;
	if (keep_geometry)
	{
		// reuse the packed geometry; only the instance-tree region of the node array is rewritten
		d_nodes = static_cast<float4*>(B[rzb_ctx::kBufNodes].ptr);
		RZB_CUDA(ctx, cudaMemsetAsync(d_nodes + 2 * size_t(top_base), 0, size_t(top_capacity) * 32, ctx->stream));
	}
	else
	{
		const MeshEntry* d_table = nullptr;
		// (triangles, raw mesh nodes and the triangle index map went first, before the host-side tree walk)
		if ((rc = uploadTo(ctx, B[rzb_ctx::kBufMeshTable], table.data(), table.size(), &d_table))) return rc;
		if ((rc = ensureBuf(ctx, B[rzb_ctx::kBufNodes], total_nodes * 32))) return rc;
		d_nodes = static_cast<float4*>(B[rzb_ctx::kBufNodes].ptr);
		RZB_CUDA(ctx, cudaMemsetAsync(d_nodes, 0, total_nodes * 32, ctx->stream));
		if (s->mesh_node_count)
		{
			k_pack_mesh_nodes<<<(s->mesh_node_count + 255) / 256, 256, 0, ctx->stream>>>(d_mesh_nodes_raw, s->mesh_node_count,
				d_table, uint32_t(table.size()), d_nodes);
			ctx->launches += 1;
		}
		if (!s->tri_host_index)
		{
			if ((rc = ensureBuf(ctx, B[rzb_ctx::kBufTriHost], size_t(s->triangle_count) * 4))) return rc;
			if (s->triangle_count)
				k_iota<<<(s->triangle_count + 255) / 256, 256, 0, ctx->stream>>>(static_cast<uint32_t*>(B[rzb_ctx::kBufTriHost].ptr), s->triangle_count);
		}
	}
	if (!keep_geometry && ctx->geom_wide)
	{
		const WideNode* d_wide = nullptr;
		if ((rc = uploadTo(ctx, B[rzb_ctx::kBufNodes4], wide_nodes.data(), wide_nodes.size(), &d_wide))) return rc;
	}
	if (top_node_count)
		RZB_CUDA(ctx, cudaMemcpyAsync(d_nodes + 2 * size_t(top_base), top_nodes.data(), top_nodes.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
	RZB_CUDA(ctx, cudaGetLastError());
	sc.nodes = d_nodes;
	sc.tri_hot = static_cast<const float4*>(B[rzb_ctx::kBufHot].ptr);
	sc.tri_cold = static_cast<const float4*>(B[rzb_ctx::kBufCold].ptr);
	sc.tri_host_index = static_cast<const uint32_t*>(B[rzb_ctx::kBufTriHost].ptr);
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufInstances], insts.data(), insts.size(), &sc.instances))) return rc;
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufInstHost], inst_host.data(), inst_host.size(), &sc.inst_host_index))) return rc;
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufInstMats], s->instance_materials, s->instance_material_count, &sc.inst_materials))) return rc;
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufMaterials], mats.data(), mats.size(), &sc.materials))) return rc;
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufMaps], maps.data(), maps.size(), &sc.maps))) return rc;
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufDirect], s->direct_lights, s->direct_light_count, &sc.direct_lights))) return rc;
	if ((rc = uploadTo(ctx, B[rzb_ctx::kBufSpot], s->spot_lights, s->spot_light_count, &sc.spot_lights))) return rc;
	std::vector<uint32_t> inst_root4;
	if (ctx->geom_wide)
	{
		inst_root4.resize(s->instance_count, kWideEmpty);
		for (uint32_t i = 0; i < s->instance_count; ++i)
			if (s->instances[i].mesh != RZB_NO_INDEX && s->instances[i].mesh < ctx->geom_mesh_root4.size())
				inst_root4[i] = ctx->geom_mesh_root4[s->instances[i].mesh];
		if ((rc = uploadTo(ctx, B[rzb_ctx::kBufInstRoot4], inst_root4.data(), inst_root4.size(), &sc.inst_root4))) return rc;
		sc.nodes4 = static_cast<const float4*>(B[rzb_ctx::kBufNodes4].ptr);
	}
	sc.top_root = top_base;
	sc.instance_count = s->instance_count;
	sc.material_count = s->material_count;
	sc.world_material = s->material_count;
	sc.default_material = s->default_material;
	sc.direct_light_count = s->direct_light_count;
	sc.spot_light_count = s->spot_light_count;
	sc.flags = ctx->cfg.flags;
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host staging vectors die at return; the caller may reuse its arrays
	ctx->sc = sc;
	if (!keep_geometry)
	{
		ctx->geom_valid = true;
		ctx->geom_mesh_base = mesh_base;
		ctx->geom_top_base = top_base;
		ctx->geom_top_capacity = top_capacity;
		ctx->geom_triangle_count = s->triangle_count;
		ctx->geom_own_trees = (s->flags & RZB_SCENE_OWN_TREES) != 0u;
	}
	ctx->own_trees = ctx->geom_own_trees;
	ctx->has_scene = true;
	// world bounds (root of the instance tree) for the Morton cells of the ray-order keys: cubic cells
	if (top_node_count)
	{
		const rzb_node& root = s->instance_nodes[0];
		float extent = 0.0f;
		for (int k = 0; k < 3; ++k)
		{
			ctx->sort_min[k] = root.bb_min[k];
			extent = std::max(extent, root.bb_max[k] - root.bb_min[k]);
		}
		ctx->sort_extent = extent > 0.0f && std::isfinite(extent) ? extent : 0.0f;
	}
	ctx->order_valid = false;
	return RZB_OK;
}

extern "C" int rzb_set_camera(rzb_ctx* ctx, const rzb_camera* camera)
{
	if (!ctx || !camera) return fail(ctx, RZB_ERR_INVALID, "rzb_set_camera: NULL argument");
	if (camera->width == 0 || camera->height == 0 || camera->width > 65536u || camera->height > 65536u ||
		uint64_t(camera->width) * camera->height > (1ull << 28))
		return fail(ctx, RZB_ERR_INVALID, "rzb_set_camera: bad resolution (1..65536 per side, at most 2^28 pixels)");
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	const bool resized = !ctx->has_camera || camera->width != ctx->cam.width || camera->height != ctx->cam.height;
	ctx->cam = *camera;
	ctx->has_camera = true;
	if (resized)
	{
		const int rc = prepareFrame(ctx);
		if (rc) return rc;
	}
	else ctx->frame.cam = makeDeviceCamera(ctx->cam);
	ctx->frame_ready = false;
	return RZB_OK;
}

extern "C" int rzb_set_rows(rzb_ctx* ctx, uint32_t row_begin, uint32_t row_end)
{
	if (!ctx) return RZB_ERR_INVALID;
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_set_rows: no camera");
	if (row_begin > row_end || row_end > ctx->cam.height) return fail(ctx, RZB_ERR_INVALID, "rzb_set_rows: bad row range");
	ctx->row_begin = row_begin;
	ctx->row_end = row_end;
	ctx->frame_ready = false; // the next render starts from a reset
	return RZB_OK;
}

extern "C" int rzb_set_row_interleave(rzb_ctx* ctx, uint32_t index, uint32_t count)
{
	if (!ctx) return RZB_ERR_INVALID;
	if (count == 0 || index >= count) return fail(ctx, RZB_ERR_INVALID, "rzb_set_row_interleave: need index < count");
	ctx->il_index = index;
	ctx->il_count = count;
	ctx->frame_ready = false;
	return RZB_OK;
}

extern "C" int rzb_set_config(rzb_ctx* ctx, const rzb_config* config)
{
	if (!ctx || !config) return fail(ctx, RZB_ERR_INVALID, "rzb_set_config: NULL argument");
	if (config->max_depth == 0 || config->max_depth > 255) return fail(ctx, RZB_ERR_INVALID, "rzb_set_config: max_depth must be 1..255");
	ctx->cfg = *config;
	ctx->sc.flags = config->flags;
	return RZB_OK;
}

extern "C" int rzb_reset(rzb_ctx* ctx)
{
	if (!ctx) return RZB_ERR_INVALID;
	if (!ctx->has_scene || !ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_reset: scene and camera must be set first");
	DeviceGuard guard(ctx->device);
	DFrame& f = ctx->frame;
	f.cam = makeDeviceCamera(ctx->cam);
	f.counters = ctx->d_counters;
	applyRows(ctx);
	const size_t n_pixels = size_t(ctx->cam.width) * ctx->cam.height;
	// temporal reprojection: keep the frame that is being replaced (Camera::swapHistoryIdx, cuda_camera.cuh:152)
	const bool reproject = (ctx->cfg.flags & RZB_FLAG_TEMPORAL_REPROJECTION) != 0u && ctx->cam.temporal_blend > 0.0f;
	if (reproject && ctx->passes > 0 && ctx->frame_cam.width == ctx->cam.width && ctx->frame_cam.height == ctx->cam.height)
	{
		if (ctx->prev_pixels != n_pixels)
		{
			if (ctx->d_prev_accum) cudaFree(ctx->d_prev_accum);
			if (ctx->d_prev_depth) cudaFree(ctx->d_prev_depth);
			ctx->d_prev_accum = nullptr; ctx->d_prev_depth = nullptr; ctx->prev_pixels = 0;
			RZB_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_prev_accum), n_pixels * 16));
			RZB_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_prev_depth), n_pixels * 4));
			ctx->prev_pixels = n_pixels;
		}
		RZB_CUDA(ctx, cudaMemcpyAsync(ctx->d_prev_accum, f.accum, n_pixels * 16, cudaMemcpyDeviceToDevice, ctx->stream));
		RZB_CUDA(ctx, cudaMemcpyAsync(ctx->d_prev_depth, f.depth, n_pixels * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		ctx->prev_cam = ctx->frame_cam;
		ctx->has_prev = true;
	}
	else ctx->has_prev = false;
	ctx->frame_cam = ctx->cam;
	RZB_CUDA(ctx, cudaMemsetAsync(f.accum, 0, n_pixels * 16, ctx->stream)); // rows outside the band stay zero
	RZB_CUDA(ctx, cudaMemsetAsync(f.depth, 0, n_pixels * 4, ctx->stream));
	RZB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 256, ctx->stream));
	if (f.slot_end > f.slot_begin)
		k_reset<<<(f.slot_end - f.slot_begin + 127) / 128, 128, 0, ctx->stream>>>(f, ctx->sc.world_material);
	RZB_CUDA(ctx, cudaGetLastError());
	RZB_CUDA(ctx, cudaMemsetAsync(ctx->d_work, 0, 128, ctx->stream));
	ctx->counted_segments = 0;
	ctx->launches += 1;
	ctx->passes = 0;
	ctx->order_valid = false;
	ctx->frame_ready = true;
	return RZB_OK;
}

extern "C" int rzb_render(rzb_ctx* ctx, uint32_t passes)
{
	if (!ctx) return RZB_ERR_INVALID;
	if (!ctx->has_scene || !ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_render: scene and camera must be set first");
	NvtxRange nvtx("rzb_render");
	DeviceGuard guard(ctx->device);
	if (!ctx->frame_ready)
	{
		const int rc = rzb_reset(ctx);
		if (rc) return rc;
	}
	int rc = ensureShadowQueue(ctx);
	if (rc) return rc;
	DFrame& f = ctx->frame;
	if (f.slot_end <= f.slot_begin) return RZB_OK; // empty row band
	f.counters = ctx->d_counters;
	f.max_depth = ctx->cfg.max_depth;
	f.direct_samples = ctx->cfg.direct_light_samples;
	f.spot_samples = ctx->cfg.spot_light_samples;
	f.inv_pdf_direct = f.direct_samples ? float(ctx->sc.direct_light_count) / float(f.direct_samples) : 0.0f;
	f.inv_pdf_spot = f.spot_samples ? float(ctx->sc.spot_light_count) / float(f.spot_samples) : 0.0f;
	f.seed = ctx->cfg.seed;
	const bool lights = (ctx->sc.direct_light_count && f.direct_samples) || (ctx->sc.spot_light_count && f.spot_samples);
	const bool count = (ctx->cfg.flags & RZB_FLAG_COUNT_WORK) != 0u;
	const bool fast = ctx->own_trees;
	f.work = ctx->d_work;
	f.prev_accum = ctx->d_prev_accum;
	f.prev_depth = ctx->d_prev_depth;
	f.prev_cam = makeDeviceCamera(ctx->prev_cam);
	f.reproject_blend = ctx->has_prev ? ctx->cam.temporal_blend : 0.0f;
	// per-stage device timing: up to 256 passes of this call are bracketed by events (4 per sampled pass)
	// With the shadow kernel overlapped (default) a stage's events bracket kernels that share the GPU, so the figures are
	// indicative only: every 8th pass is sampled there (five event records per sampled pass cost 13 us of a 1.2 ms pass:
	// 1.216 -> 1.203 ms per pass in a 20-pass call). In stream order (RZB_FLAG_SERIAL_STAGES, counting kernels,
	// RZB200_OVERLAP=0) every pass is sampled. RZB200_STAGE_STRIDE overrides the stride.
	const bool overlapped = ctx->overlap && lights && !count && !ctx->debug_sync && !(ctx->cfg.flags & RZB_FLAG_SERIAL_STAGES);
	const uint32_t stride = std::max((passes + 255u) / 256u, ctx->stage_stride ? ctx->stage_stride : (overlapped ? 8u : 1u));
	const uint32_t n_sampled = passes ? (passes + stride - 1u) / stride : 0u;
	const uint32_t n_slots = f.slot_end - f.slot_begin;
	// ---- ray ordering set-up: bin tables = [camera groups | bounce bins | 1 bin for slots without a pixel | shadow bins]
	size_t scan_temp = 0;
	uint32_t n_bins = 0;
	if (ctx->sort_enabled)
	{
		const uint32_t dir_bins_log2 = ctx->sort_dir_bits ? 2u * ctx->sort_dir_bits : 3u;
		f.sort_bits = ctx->sort_bits;
		f.sort_dir_bits = ctx->sort_dir_bits;
		f.sort_camera_bins = (n_slots + 31u) / 32u;
		f.sort_bounce_bins = 1u << (3u * ctx->sort_bits + dir_bins_log2);
		f.sort_shadow_base = f.sort_camera_bins + f.sort_bounce_bins + 1u;
		f.sort_shadow_bits = ctx->sort_shadow_bits;
		f.order_reversed = ctx->order_reversed ? 1u : 0u;
		f.sort_dir_major = ctx->sort_dir_major ? 1u : 0u;
		const uint32_t shadow_bins = (lights && ctx->sort_shadow) ? (1u << (3u * ctx->sort_shadow_bits + 2u)) : 0u;
		n_bins = f.sort_shadow_base + shadow_bins;
		for (int k = 0; k < 3; ++k) f.sort_min[k] = ctx->sort_min[k];
		f.sort_scale = ctx->sort_extent > 0.0f ? 1.0f / ctx->sort_extent : 0.0f; // to [0, 1); the kernels scale by their cell count
		scan_temp = scanTempBytes(n_bins + 1u);
		const bool grow = ctx->sort_buf[rzb_ctx::kSortKeys].bytes < size_t(n_slots) * 4 || ctx->sort_buf[rzb_ctx::kSortBins].bytes < size_t(n_bins + 1u) * 4;
		if ((rc = ensureBuf(ctx, ctx->sort_buf[rzb_ctx::kSortTemp], scan_temp))) return rc;
		for (int k : {int(rzb_ctx::kSortKeys), int(rzb_ctx::kSortRank), int(rzb_ctx::kSortOrder)})
			if ((rc = ensureBuf(ctx, ctx->sort_buf[k], size_t(n_slots) * 4))) return rc;
		for (int k : {int(rzb_ctx::kSortBins), int(rzb_ctx::kSortOffsets)})
			if ((rc = ensureBuf(ctx, ctx->sort_buf[k], size_t(n_bins + 1u) * 4))) return rc;
		if (shadow_bins)
			for (int k : {int(rzb_ctx::kSortShKeys), int(rzb_ctx::kSortShRank), int(rzb_ctx::kSortShOrder)})
				if ((rc = ensureBuf(ctx, ctx->sort_buf[k], size_t(f.shadow_capacity) * 4))) return rc;
		if (grow) ctx->order_valid = false;
		f.sort_keys = static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortKeys].ptr);
		f.sort_rank = static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortRank].ptr);
		f.sort_bin_count = static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortBins].ptr);
		f.sh_keys = shadow_bins ? static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortShKeys].ptr) : nullptr;
		f.sh_rank = shadow_bins ? static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortShRank].ptr) : nullptr;
		RZB_CUDA(ctx, cudaMemsetAsync(f.sort_bin_count, 0, size_t(n_bins + 1u) * 4, ctx->stream));
	}
	else { f.sort_keys = nullptr; f.sort_rank = nullptr; f.sort_bin_count = nullptr; f.sh_keys = nullptr; f.sh_rank = nullptr; f.order = nullptr; }
	f.sh_order = nullptr;
	while (ctx->ev_stage.size() < size_t(n_sampled) * 5)
	{
		cudaEvent_t ev = nullptr;
		RZB_CUDA(ctx, cudaEventCreate(&ev));
		ctx->ev_stage.push_back(ev);
	}
	ctx->sampled_passes = 0;
	ctx->last_had_shadow = lights;
	RZB_CUDA(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream));
	for (uint32_t p = 0; p < passes; ++p)
	{
		f.pass_index = uint32_t(ctx->passes);
		const bool timed = (p % stride) == 0u;
		cudaEvent_t* ev = timed ? &ctx->ev_stage[size_t(ctx->sampled_passes) * 5] : nullptr;
		// (overlap: the shadow kernel of the previous pass may still be reading its counters -- alternate between two sets)
		f.counters = ctx->d_counters + ((ctx->overlap && (ctx->passes & 1u)) ? 20 : 0);
		RZB_CUDA(ctx, cudaMemsetAsync(f.counters, 0, 12, ctx->stream));
		f.order = (ctx->sort_enabled && ctx->order_valid) ? static_cast<const uint32_t*>(ctx->sort_buf[rzb_ctx::kSortOrder].ptr) : nullptr;
		if (timed) cudaEventRecord(ev[0], ctx->stream);
		if (fast && ctx->geom_wide)
		{
			if (count) k_trace_paths<true, true, false, true><<<ctx->wide_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f);
			else k_trace_paths<false, true, false, true><<<ctx->wide_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f);
		}
		else if (ctx->trace_mr)
		{
			const MrPathKernel kernel = reinterpret_cast<MrPathKernel>(const_cast<void*>(mrPathKernel(ctx->mr_k, ctx->mr_steps, count, fast)));
			kernel<<<ctx->mr_grid[fast ? 1 : 0], kMrBlock, ctx->mr_smem, ctx->stream>>>(ctx->sc, f);
		}
		else if (ctx->trace_mode == 1 && !count) { if (fast) k_trace_paths<false, true, 1><<<ctx->trace_grid_fast, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f); else k_trace_paths<false, false, 1><<<ctx->trace_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f); }
		else if (count) { if (fast) k_trace_paths<true, true><<<ctx->trace_grid_fast, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f); else k_trace_paths<true, false><<<ctx->trace_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f); }
		else { if (fast) k_trace_paths<false, true><<<ctx->trace_grid_fast, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f); else k_trace_paths<false, false><<<ctx->trace_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc, f); }
		if (ctx->debug_sync)
		{
			const cudaError_t e = cudaStreamSynchronize(ctx->stream);
			if (e != cudaSuccess) return cudaFail(ctx, e, ("k_trace_paths, pass " + std::to_string(ctx->passes)).c_str());
		}
		if (ctx->shadow_pending)
		{
			// the previous pass's shadow kernel adds to the accumulator and reads the shadow queue k_shade refills
			RZB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
			ctx->shadow_pending = false;
		}
		if (f.pass_index == 0u && f.reproject_blend > 0.0f)
		{
			k_reproject<<<(f.slot_end - f.slot_begin + 127) / 128, 128, 0, ctx->stream>>>(f);
			ctx->launches += 1;
		}
		if (timed) cudaEventRecord(ev[1], ctx->stream);
		k_shade<<<(f.slot_end - f.slot_begin + 127) / 128, 128, 0, ctx->stream>>>(ctx->sc, f);
		if (ctx->debug_sync)
		{
			const cudaError_t e = cudaStreamSynchronize(ctx->stream);
			if (e != cudaSuccess) return cudaFail(ctx, e, ("k_shade, pass " + std::to_string(ctx->passes)).c_str());
		}
		if (timed) cudaEventRecord(ev[2], ctx->stream);
		ctx->launches += 2;
		if (ctx->sort_enabled)
		{
			// bins -> offsets (one prefix sum over both tables), then scatter: the order of the NEXT pass's closest-hit
			// queries and of THIS pass's shadow queries
			uint32_t* offsets = static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortOffsets].ptr);
			RZB_CUDA(ctx, exclusiveScan(ctx->sort_buf[rzb_ctx::kSortTemp].ptr, scan_temp, f.sort_bin_count, offsets, n_bins + 1u, ctx->stream));
			RZB_CUDA(ctx, cudaMemsetAsync(f.sort_bin_count, 0, size_t(n_bins + 1u) * 4, ctx->stream));
			uint32_t* sh_order = f.sh_keys ? static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortShOrder].ptr) : nullptr;
			const uint32_t scatter_n = std::max(n_slots, f.sh_keys ? f.shadow_capacity : 0u);
			k_scatter_order<<<(scatter_n + 255) / 256, 256, 0, ctx->stream>>>(f, offsets, static_cast<uint32_t*>(ctx->sort_buf[rzb_ctx::kSortOrder].ptr), sh_order);
			f.sh_order = sh_order;
			ctx->order_valid = true;
			ctx->launches += 3;
		}
		if (timed) cudaEventRecord(ev[4], ctx->stream);
		if (lights)
		{
			// (wide trees: the any-hit walk stays on the binary trees, which are on the device too -- measured faster there:
			// 64 registers / 8 blocks per SM against 72 / 7, and an any-hit walk gains nothing from nearest-of-four ordering)
			const bool side = ctx->overlap && !count && !ctx->debug_sync && !(ctx->cfg.flags & RZB_FLAG_SERIAL_STAGES);
			cudaStream_t sh_stream = side ? ctx->stream2 : ctx->stream;
			if (side)
			{
				RZB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
				RZB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
			}
			if (count) k_trace_shadow<true><<<ctx->shadow_grid, kTraceBlock, 0, sh_stream>>>(ctx->sc, f);
			else k_trace_shadow<false><<<ctx->shadow_grid, kTraceBlock, 0, sh_stream>>>(ctx->sc, f);
			if (side)
			{
				if (timed) cudaEventRecord(ev[3], ctx->stream2);
				RZB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
				ctx->shadow_pending = true;
			}
			ctx->launches += 1;
			if (ctx->debug_sync)
			{
				const cudaError_t e = cudaStreamSynchronize(ctx->stream);
				if (e != cudaSuccess) return cudaFail(ctx, e, ("k_trace_shadow, pass " + std::to_string(ctx->passes)).c_str());
			}
		}
		if (timed)
		{
			if (!ctx->shadow_pending) cudaEventRecord(ev[3], ctx->stream);
			ctx->sampled_passes += 1;
		}
		ctx->passes += 1;
		if (count) ctx->counted_segments += bandPixels(ctx);
	}
	if (ctx->shadow_pending)
	{
		RZB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
		ctx->shadow_pending = false;
	}
	RZB_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->stream));
	RZB_CUDA(ctx, cudaGetLastError());
	return RZB_OK;
}

extern "C" int rzb_synchronize(rzb_ctx* ctx)
{
	if (!ctx) return RZB_ERR_INVALID;
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RZB_OK;
}

namespace
{
	int tonemapAndCopy(rzb_ctx* ctx, const PeerList& peers, uint8_t* rgba8, float* depth, bool sync = true)
	{
		NvtxRange nvtx("rzb_resolve");
		const uint32_t n = ctx->cam.width * ctx->cam.height;
		if (rgba8)
		{
			// ComputeFinalColor multiplies by aperture area, exposure time and 1e5 one after the other
			const float area = ctx->cam.aperture * ctx->cam.aperture * 3.14159265358979323846f;
			k_tonemap<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->frame.accum, peers, ctx->d_rgba, n, area, ctx->cam.exposure_time);
			ctx->launches += 1;
			RZB_CUDA(ctx, cudaGetLastError());
			RZB_CUDA(ctx, cudaMemcpyAsync(rgba8, ctx->d_rgba, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
		}
		if (depth) RZB_CUDA(ctx, cudaMemcpyAsync(depth, ctx->frame.depth, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (sync) RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		return RZB_OK;
	}
}

extern "C" int rzb_resolve(rzb_ctx* ctx, uint8_t* rgba8, float* depth, uint64_t* ray_count)
{
	if (!ctx) return RZB_ERR_INVALID;
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_resolve: no camera");
	DeviceGuard guard(ctx->device);
	PeerList peers{};
	const int rc = tonemapAndCopy(ctx, peers, rgba8, depth);
	if (rc) return rc;
	if (ray_count) *ray_count = ctx->passes * bandPixels(ctx);
	return RZB_OK;
}

extern "C" int rzb_resolve_async(rzb_ctx* ctx, uint32_t slot, uint8_t* rgba8_pinned, float* depth_pinned, uint64_t* ray_count)
{
	if (!ctx || slot > 1u) return fail(ctx, RZB_ERR_INVALID, "rzb_resolve_async: bad argument");
	if (!ctx->has_camera || !ctx->has_scene) return fail(ctx, RZB_ERR_STATE, "rzb_resolve_async: scene and camera must be set first");
	DeviceGuard guard(ctx->device);
	PeerList peers{};
	const int rc = tonemapAndCopy(ctx, peers, rgba8_pinned, depth_pinned, false);
	if (rc) return rc;
	// the ray-cast pick travels with the frame (rayCast kernel, cuda_render_kernel.cu:130-144)
	const uint32_t px = std::min(ctx->cam.raycast_pixel[0], ctx->cam.width - 1u);
	const uint32_t py = std::min(ctx->cam.raycast_pixel[1], ctx->cam.height - 1u);
	uint32_t* d_out = ctx->d_counters + 44 + 2 * slot;
	k_raycast<<<1, 32, 0, ctx->stream>>>(ctx->sc, makeDeviceCamera(ctx->cam), ctx->frame.depth, px, py, d_out);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	RZB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pick + 2 * slot, d_out, 8, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaEventRecord(ctx->ev_resolve[slot], ctx->stream));
	if (ray_count) *ray_count = ctx->passes * bandPixels(ctx);
	return RZB_OK;
}

extern "C" int rzb_resolve_wait(rzb_ctx* ctx, uint32_t slot, uint32_t* instance, uint32_t* material_slot)
{
	if (!ctx || slot > 1u) return fail(ctx, RZB_ERR_INVALID, "rzb_resolve_wait: bad argument");
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaEventSynchronize(ctx->ev_resolve[slot]));
	if (instance) *instance = ctx->h_pick[2 * slot];
	if (material_slot) *material_slot = ctx->h_pick[2 * slot + 1];
	return RZB_OK;
}

extern "C" int rzb_host_alloc(size_t bytes, void** out)
{
	if (!out || bytes == 0) return RZB_ERR_INVALID;
	*out = nullptr;
	return cudaMallocHost(out, bytes) == cudaSuccess ? RZB_OK : RZB_ERR_NOMEM;
}

extern "C" int rzb_host_free(void* p)
{
	if (!p) return RZB_OK;
	return cudaFreeHost(p) == cudaSuccess ? RZB_OK : RZB_ERR_CUDA;
}

extern "C" int rzb_resolve_peers(rzb_ctx* ctx, rzb_ctx* const* peers_in, uint32_t n_peers,
	uint8_t* rgba8, float* depth, uint64_t* ray_count)
{
	if (!ctx || (n_peers && !peers_in) || n_peers > 8) return fail(ctx, RZB_ERR_INVALID, "rzb_resolve_peers: bad arguments");
	DeviceGuard guard(ctx->device);
	PeerList peers{};
	uint64_t rays = ctx->passes * bandPixels(ctx);
	for (uint32_t i = 0; i < n_peers; ++i)
	{
		rzb_ctx* p = peers_in[i];
		if (!p || p->cam.width != ctx->cam.width || p->cam.height != ctx->cam.height)
			return fail(ctx, RZB_ERR_INVALID, "rzb_resolve_peers: peer resolution differs");
		if (p->device != ctx->device)
		{
			int can = 0;
			RZB_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, p->device));
			if (!can) return fail(ctx, RZB_ERR_CUDA, "rzb_resolve_peers: no peer access between devices");
			const cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cudaFail(ctx, e, "cudaDeviceEnablePeerAccess");
			cudaGetLastError();
		}
		{
			DeviceGuard pg(p->device);
			RZB_CUDA(ctx, cudaStreamSynchronize(p->stream));
		}
		peers.accum[i] = p->frame.accum;
		rays += p->passes * bandPixels(p);
	}
	peers.count = n_peers;
	const int rc = tonemapAndCopy(ctx, peers, rgba8, depth);
	if (rc) return rc;
	if (ray_count) *ray_count = rays;
	return RZB_OK;
}

extern "C" int rzb_read_accum(rzb_ctx* ctx, float* rgba_f32)
{
	if (!ctx || !rgba_f32) return fail(ctx, RZB_ERR_INVALID, "rzb_read_accum: NULL argument");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_read_accum: no camera");
	DeviceGuard guard(ctx->device);
	const size_t n = size_t(ctx->cam.width) * ctx->cam.height;
	RZB_CUDA(ctx, cudaMemcpyAsync(rgba_f32, ctx->frame.accum, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RZB_OK;
}

extern "C" int rzb_mean_samples(rzb_ctx* ctx, double* mean_out)
{
	if (!ctx || !mean_out) return fail(ctx, RZB_ERR_INVALID, "rzb_mean_samples: NULL argument");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_mean_samples: no camera");
	DeviceGuard guard(ctx->device);
	const uint32_t n = ctx->cam.width * ctx->cam.height;
	double* d_sum = reinterpret_cast<double*>(ctx->d_work + 15);
	RZB_CUDA(ctx, cudaMemsetAsync(d_sum, 0, 8, ctx->stream));
	k_sum_alpha<<<std::min<uint32_t>((n + 255u) / 256u, uint32_t(ctx->sm_count) * 8u), 256, 0, ctx->stream>>>(ctx->frame.accum, n, d_sum);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	double h = 0.0;
	RZB_CUDA(ctx, cudaMemcpyAsync(&h, d_sum, 8, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	const uint64_t band = bandPixels(ctx);
	*mean_out = band ? h / double(band) : 0.0;
	return RZB_OK;
}

extern "C" int rzb_accum_device_ptr(rzb_ctx* ctx, void** device_ptr, size_t* bytes)
{
	if (!ctx || !device_ptr) return fail(ctx, RZB_ERR_INVALID, "rzb_accum_device_ptr: NULL argument");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_accum_device_ptr: no camera");
	*device_ptr = ctx->frame.accum;
	if (bytes) *bytes = size_t(ctx->cam.width) * ctx->cam.height * 16;
	return RZB_OK;
}

extern "C" int rzb_accum_add_device(rzb_ctx* ctx, const void* device_rgba_f32, size_t pixel_count)
{
	if (!ctx || !device_rgba_f32) return fail(ctx, RZB_ERR_INVALID, "rzb_accum_add_device: NULL argument");
	if (!ctx->has_camera || pixel_count > size_t(ctx->cam.width) * ctx->cam.height)
		return fail(ctx, RZB_ERR_INVALID, "rzb_accum_add_device: too many pixels");
	DeviceGuard guard(ctx->device);
	const uint32_t n = uint32_t(pixel_count);
	k_accum_add<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->frame.accum, static_cast<const float4*>(device_rgba_f32), n);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	return RZB_OK;
}

extern "C" int rzb_get_render_stats(rzb_ctx* ctx, rzb_render_stats* out)
{
	if (!ctx || !out) return RZB_ERR_INVALID;
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	std::memset(out, 0, sizeof(*out));
	out->passes = ctx->passes;
	out->ray_count = ctx->passes * bandPixels(ctx);
	out->kernel_launches = ctx->launches;
	if (ctx->passes)
	{
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess) ctx->last_render_ms = ms;
		double t = 0.0, sh = 0.0, sd = 0.0, so = 0.0;
		for (uint32_t i = 0; i < ctx->sampled_passes; ++i)
		{
			const cudaEvent_t* ev = &ctx->ev_stage[size_t(i) * 5];
			if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) t += ms;
			if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) sh += ms;
			if (cudaEventElapsedTime(&ms, ev[2], ev[4]) == cudaSuccess) so += ms;
			if (ctx->last_had_shadow && cudaEventElapsedTime(&ms, ev[4], ev[3]) == cudaSuccess) sd += ms;
		}
		if (ctx->sampled_passes)
		{
			ctx->last_sort_ms = float(so / ctx->sampled_passes);
			ctx->last_trace_ms = float(t / ctx->sampled_passes);
			ctx->last_shade_ms = float(sh / ctx->sampled_passes);
			ctx->last_shadow_ms = float(sd / ctx->sampled_passes);
		}
		cudaGetLastError();
	}
	uint32_t counters[32] = {};
	RZB_CUDA(ctx, cudaMemcpy(counters, ctx->d_counters, 128, cudaMemcpyDeviceToHost));
	out->shadow_rays = counters[(ctx->overlap && ctx->passes && ((ctx->passes - 1u) & 1u)) ? 21 : 1]; // of the last pass
	out->last_render_ms = ctx->last_render_ms;
	out->last_trace_ms = ctx->last_trace_ms;
	out->last_shade_ms = ctx->last_shade_ms;
	out->last_shadow_ms = ctx->last_shadow_ms;
	out->last_sort_ms = ctx->last_sort_ms;
	out->last_exchange_ms = ctx->last_exchange_ms;
	return RZB_OK;
}

extern "C" int rzb_get_work_counters(rzb_ctx* ctx, rzb_work_counters* out)
{
	if (!ctx || !out) return RZB_ERR_INVALID;
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	unsigned long long h[16] = {};
	RZB_CUDA(ctx, cudaMemcpy(h, ctx->d_work, sizeof(h), cudaMemcpyDeviceToHost));
	out->closest_top_nodes = h[0]; out->closest_instances = h[1]; out->closest_mesh_nodes = h[2]; out->closest_triangles = h[3];
	out->shadow_top_nodes = h[4]; out->shadow_instances = h[5]; out->shadow_mesh_nodes = h[6]; out->shadow_triangles = h[7];
	out->shadow_rays = h[8];
	out->segments = ctx->counted_segments;
	out->invalid_rays = h[9];
	out->closest_lane_work = h[10]; out->closest_batch_work = h[11];
	out->shadow_lane_work = h[12]; out->shadow_batch_work = h[13];
	return RZB_OK;
}

extern "C" int rzb_accum_ipc_handle(rzb_ctx* ctx, void* handle_out)
{
	if (!ctx || !handle_out) return fail(ctx, RZB_ERR_INVALID, "rzb_accum_ipc_handle: NULL argument");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_accum_ipc_handle: no camera");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	DeviceGuard guard(ctx->device);
	cudaIpcMemHandle_t h;
	RZB_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->frame.accum));
	std::memcpy(handle_out, &h, 64);
	return RZB_OK;
}

extern "C" int rzb_resolve_ipc(rzb_ctx* ctx, const void* handles, uint32_t n_peers, uint8_t* rgba8, float* depth)
{
	if (!ctx || (n_peers && !handles) || n_peers > 8) return fail(ctx, RZB_ERR_INVALID, "rzb_resolve_ipc: bad arguments");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_resolve_ipc: no camera");
	DeviceGuard guard(ctx->device);
	PeerList peers{};
	for (uint32_t i = 0; i < n_peers; ++i)
	{
		const std::string key(static_cast<const char*>(handles) + size_t(i) * 64, 64);
		void* mapped = nullptr;
		for (auto& h : ctx->ipc_open) if (h.first == key) mapped = h.second;
		if (!mapped)
		{
			cudaIpcMemHandle_t h;
			std::memcpy(&h, key.data(), 64);
			RZB_CUDA(ctx, cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
			ctx->ipc_open.emplace_back(key, mapped);
		}
		peers.accum[i] = static_cast<const float4*>(mapped);
	}
	peers.count = n_peers;
	return tonemapAndCopy(ctx, peers, rgba8, depth);
}

namespace
{
	int openIpc(rzb_ctx* ctx, const char* handle64, void** mapped_out)
	{
		const std::string key(handle64, 64);
		for (auto& h : ctx->ipc_open)
			if (h.first == key) { *mapped_out = h.second; return RZB_OK; }
		cudaIpcMemHandle_t h;
		std::memcpy(&h, key.data(), 64);
		void* mapped = nullptr;
		RZB_CUDA(ctx, cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
		ctx->ipc_open.emplace_back(key, mapped);
		*mapped_out = mapped;
		return RZB_OK;
	}
	int ensureExchange(rzb_ctx* ctx)
	{
		const size_t n = size_t(ctx->cam.width) * ctx->cam.height;
		if (ctx->d_exchange && ctx->exchange_pixels == n) return RZB_OK;
		if (ctx->d_exchange) cudaFree(ctx->d_exchange);
		ctx->d_exchange = nullptr;
		ctx->exchange_pixels = 0;
		RZB_CUDA(ctx, cudaMalloc(&ctx->d_exchange, sizeof(ExchangeHeader) + n * 4));
		RZB_CUDA(ctx, cudaMemset(ctx->d_exchange, 0, sizeof(ExchangeHeader)));
		ctx->exchange_pixels = n;
		ctx->exchange_epoch = 0;
		return RZB_OK;
	}
}

extern "C" int rzb_exchange_ipc_handle(rzb_ctx* ctx, void* handle_out)
{
	if (!ctx || !handle_out) return fail(ctx, RZB_ERR_INVALID, "rzb_exchange_ipc_handle: NULL argument");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_exchange_ipc_handle: no camera");
	DeviceGuard guard(ctx->device);
	const int rc = ensureExchange(ctx);
	if (rc) return rc;
	cudaIpcMemHandle_t h;
	RZB_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_exchange));
	std::memcpy(handle_out, &h, 64);
	return RZB_OK;
}

extern "C" int rzb_resolve_sliced(rzb_ctx* ctx, uint32_t rank, uint32_t world, const void* accum_handles,
	const void* exchange_handles, uint8_t* rgba8_pinned, float* depth_pinned)
{
	if (!ctx || world == 0 || world > 8 || rank >= world || (world > 1 && (!accum_handles || !exchange_handles)))
		return fail(ctx, RZB_ERR_INVALID, "rzb_resolve_sliced: bad arguments");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_resolve_sliced: no camera");
	DeviceGuard guard(ctx->device);
	int rc = ensureExchange(ctx);
	if (rc) return rc;
	SlicedArgs a{};
	for (uint32_t r = 0; r < world; ++r)
	{
		if (r == rank)
		{
			a.accum[r] = ctx->frame.accum;
			a.header[r] = static_cast<ExchangeHeader*>(ctx->d_exchange);
			continue;
		}
		void* m = nullptr;
		if ((rc = openIpc(ctx, static_cast<const char*>(accum_handles) + size_t(r) * 64, &m))) return rc;
		a.accum[r] = static_cast<const float4*>(m);
		if ((rc = openIpc(ctx, static_cast<const char*>(exchange_handles) + size_t(r) * 64, &m))) return rc;
		a.header[r] = static_cast<ExchangeHeader*>(m);
	}
	const uint32_t n = ctx->cam.width * ctx->cam.height;
	a.root_rgba = reinterpret_cast<uchar4*>(reinterpret_cast<char*>(a.header[0]) + sizeof(ExchangeHeader));
	a.rank = rank; a.world = world;
	a.epoch = ++ctx->exchange_epoch;
	// slices of whole 32-pixel groups, so that every 128-byte line of the staging image has one writer
	const uint32_t groups = (n + 31u) / 32u;
	a.begin = std::min(n, uint32_t((uint64_t(groups) * rank / world) * 32u));
	a.end = std::min(n, uint32_t((uint64_t(groups) * (rank + 1u) / world) * 32u));
	a.aperture_area = ctx->cam.aperture * ctx->cam.aperture * 3.14159265358979323846f;
	a.exposure_time = ctx->cam.exposure_time;
	a.spin_limit = 20ull * 1000ull * 1000ull * 1000ull; // ~10 s of SM clock: a peer that never arrives ends the wait
	for (cudaEvent_t& ev : ctx->ev_exchange)
		if (!ev) RZB_CUDA(ctx, cudaEventCreate(&ev));
	RZB_CUDA(ctx, cudaEventRecord(ctx->ev_exchange[0], ctx->stream));
	const uint32_t slice = a.end - a.begin;
	const int grid = int(std::max<uint32_t>(1u, std::min<uint32_t>((slice + 255u) / 256u, uint32_t(ctx->sm_count) * 4u)));
	k_resolve_sliced<<<grid, 256, 0, ctx->stream>>>(a);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	RZB_CUDA(ctx, cudaEventRecord(ctx->ev_exchange[1], ctx->stream));
	ctx->exchange_timed = true;
	if (rank == 0)
	{
		if (rgba8_pinned) RZB_CUDA(ctx, cudaMemcpyAsync(rgba8_pinned, a.root_rgba, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (depth_pinned) RZB_CUDA(ctx, cudaMemcpyAsync(depth_pinned, ctx->frame.depth, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	// completion + "a peer never arrived" travel through slot 0 of the asynchronous-resolve protocol
	if (ctx->h_pick)
		RZB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pick + 4, &static_cast<ExchangeHeader*>(ctx->d_exchange)->timed_out, 4, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaEventRecord(ctx->ev_resolve[0], ctx->stream));
	return RZB_OK;
}

extern "C" int rzb_resolve_sliced_wait(rzb_ctx* ctx, float* exchange_ms_or_null)
{
	if (!ctx) return RZB_ERR_INVALID;
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaEventSynchronize(ctx->ev_resolve[0]));
	if (ctx->exchange_timed)
	{
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, ctx->ev_exchange[0], ctx->ev_exchange[1]) == cudaSuccess) ctx->last_exchange_ms = ms;
		cudaGetLastError();
	}
	if (exchange_ms_or_null) *exchange_ms_or_null = ctx->last_exchange_ms;
	if (ctx->h_pick && ctx->h_pick[4] != 0u)
		return fail(ctx, RZB_ERR_STATE, "rzb_resolve_sliced: a peer rank never reached the exchange step (spin limit hit); the frame is incomplete");
	return RZB_OK;
}

extern "C" int rzb_timings(rzb_ctx* ctx, char* buf, size_t buf_size)
{
	if (!ctx || !buf || buf_size == 0) return RZB_ERR_INVALID;
	rzb_render_stats st{};
	const int rc = rzb_get_render_stats(ctx, &st);
	if (rc) return rc;
	std::snprintf(buf, buf_size,
		"B200 wavefront engine (device %d, %d SMs)\n"
		"passes: %llu  rays: %llu  kernel launches: %llu\n"
		"last render call: %.3f ms\n"
		"  closest-hit: %.3f ms/pass\n  shade+NEE:   %.3f ms/pass\n  shadow rays: %.3f ms/pass (%llu rays)\n",
		ctx->device, ctx->sm_count, (unsigned long long)st.passes, (unsigned long long)st.ray_count,
		(unsigned long long)st.kernel_launches, st.last_render_ms, st.last_trace_ms, st.last_shade_ms,
		st.last_shadow_ms, (unsigned long long)st.shadow_rays);
	return RZB_OK;
}

// ---------------------------------------------------------------- ray-set entry points
namespace
{
	int packRays(rzb_ctx* ctx, const float* origins, const float* directions, const float* near_far, uint32_t n)
	{
		int rc;
		if ((rc = ensureScratch(ctx, 0, size_t(n) * 16))) return rc;
		if ((rc = ensureScratch(ctx, 1, size_t(n) * 16))) return rc;
		std::vector<float4> o(n), d(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			o[i] = make_float4(origins[3 * size_t(i)], origins[3 * size_t(i) + 1], origins[3 * size_t(i) + 2], near_far[2 * size_t(i)]);
			d[i] = make_float4(directions[3 * size_t(i)], directions[3 * size_t(i) + 1], directions[3 * size_t(i) + 2], near_far[2 * size_t(i) + 1]);
		}
		// on the context's stream (a non-blocking stream has no implicit ordering with the legacy default stream); the
		// staging vectors die at return, so wait for the copies here -- the kernels that follow are stream-ordered anyway
		RZB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[0].ptr, o.data(), size_t(n) * 16, cudaMemcpyHostToDevice, ctx->stream));
		RZB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch[1].ptr, d.data(), size_t(n) * 16, cudaMemcpyHostToDevice, ctx->stream));
		RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		return RZB_OK;
	}
}

extern "C" int rzb_trace_closest_device(rzb_ctx* ctx, const void* rays_o_near, const void* rays_d_far,
	uint32_t n, void* hits_out_device, float* elapsed_ms)
{
	if (!ctx || !rays_o_near || !rays_d_far || !hits_out_device) return fail(ctx, RZB_ERR_INVALID, "rzb_trace_closest_device: NULL argument");
	if (!ctx->has_scene) return fail(ctx, RZB_ERR_STATE, "rzb_trace_closest_device: no scene");
	DeviceGuard guard(ctx->device);
	RZB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 4, ctx->stream));
	if (elapsed_ms) RZB_CUDA(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream));
	if (ctx->own_trees && ctx->geom_wide)
		k_trace_rays<false, true, true><<<ctx->rays_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
			static_cast<const float4*>(rays_o_near), static_cast<const float4*>(rays_d_far), n,
			static_cast<DHit*>(hits_out_device), ctx->d_counters + 8, nullptr);
	else if (ctx->trace_mr)
		reinterpret_cast<MrRaysKernel>(const_cast<void*>(mrRaysKernel(ctx->mr_k, ctx->mr_steps, ctx->own_trees)))<<<ctx->mr_grid[ctx->own_trees ? 1 : 0], kMrBlock, ctx->mr_smem, ctx->stream>>>(ctx->sc,
			static_cast<const float4*>(rays_o_near), static_cast<const float4*>(rays_d_far), n,
			static_cast<DHit*>(hits_out_device), ctx->d_counters + 8, nullptr);
	else
	(ctx->own_trees ? k_trace_rays<false, true> : k_trace_rays<false, false>)<<<ctx->rays_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
		static_cast<const float4*>(rays_o_near), static_cast<const float4*>(rays_d_far), n,
		static_cast<DHit*>(hits_out_device), ctx->d_counters + 8, nullptr);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	if (elapsed_ms)
	{
		RZB_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->stream));
		RZB_CUDA(ctx, cudaEventSynchronize(ctx->ev_end));
		RZB_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev_begin, ctx->ev_end));
	}
	return RZB_OK;
}

extern "C" int rzb_trace_closest_device_counted(rzb_ctx* ctx, const void* rays_o_near, const void* rays_d_far,
	uint32_t n, void* hits_out_device)
{
	if (!ctx || !rays_o_near || !rays_d_far || !hits_out_device) return fail(ctx, RZB_ERR_INVALID, "rzb_trace_closest_device_counted: NULL argument");
	if (!ctx->has_scene) return fail(ctx, RZB_ERR_STATE, "rzb_trace_closest_device_counted: no scene");
	DeviceGuard guard(ctx->device);
	unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(ctx->d_counters + 10);
	RZB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 40, ctx->stream));
	(ctx->own_trees ? (ctx->geom_wide ? k_trace_rays<true, true, true> : k_trace_rays<true, true, false>) : k_trace_rays<true, false, false>)<<<ctx->rays_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
		static_cast<const float4*>(rays_o_near), static_cast<const float4*>(rays_d_far), n,
		static_cast<DHit*>(hits_out_device), ctx->d_counters + 8, d_stats);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RZB_OK;
}

extern "C" int rzb_trace_closest(rzb_ctx* ctx, const float* origins, const float* directions,
	const float* near_far, uint32_t n, rzb_hit* hits_out, rzb_trace_stats* stats)
{
	if (!ctx || !origins || !directions || !near_far || !hits_out) return fail(ctx, RZB_ERR_INVALID, "rzb_trace_closest: NULL argument");
	if (!ctx->has_scene) return fail(ctx, RZB_ERR_STATE, "rzb_trace_closest: no scene");
	if (n == 0) return RZB_OK;
	DeviceGuard guard(ctx->device);
	int rc;
	if ((rc = packRays(ctx, origins, directions, near_far, n))) return rc;
	if ((rc = ensureScratch(ctx, 2, size_t(n) * sizeof(DHit)))) return rc;
	if ((rc = ensureScratch(ctx, 3, size_t(n) * sizeof(rzb_hit) + 64))) return rc;
	unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(ctx->d_counters + 10);
	RZB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 40, ctx->stream));
	if (stats)
		(ctx->own_trees ? (ctx->geom_wide ? k_trace_rays<true, true, true> : k_trace_rays<true, true, false>) : k_trace_rays<true, false, false>)<<<ctx->rays_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
			static_cast<const float4*>(ctx->scratch[0].ptr), static_cast<const float4*>(ctx->scratch[1].ptr), n,
			static_cast<DHit*>(ctx->scratch[2].ptr), ctx->d_counters + 8, d_stats);
	else if (ctx->own_trees && ctx->geom_wide)
		k_trace_rays<false, true, true><<<ctx->rays_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
			static_cast<const float4*>(ctx->scratch[0].ptr), static_cast<const float4*>(ctx->scratch[1].ptr), n,
			static_cast<DHit*>(ctx->scratch[2].ptr), ctx->d_counters + 8, nullptr);
	else if (ctx->trace_mr)
		reinterpret_cast<MrRaysKernel>(const_cast<void*>(mrRaysKernel(ctx->mr_k, ctx->mr_steps, ctx->own_trees)))<<<ctx->mr_grid[ctx->own_trees ? 1 : 0], kMrBlock, ctx->mr_smem, ctx->stream>>>(ctx->sc,
			static_cast<const float4*>(ctx->scratch[0].ptr), static_cast<const float4*>(ctx->scratch[1].ptr), n,
			static_cast<DHit*>(ctx->scratch[2].ptr), ctx->d_counters + 8, nullptr);
	else
		(ctx->own_trees ? k_trace_rays<false, true> : k_trace_rays<false, false>)<<<ctx->rays_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
			static_cast<const float4*>(ctx->scratch[0].ptr), static_cast<const float4*>(ctx->scratch[1].ptr), n,
			static_cast<DHit*>(ctx->scratch[2].ptr), ctx->d_counters + 8, nullptr);
	k_convert_hits<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->sc, static_cast<const DHit*>(ctx->scratch[2].ptr),
		static_cast<rzb_hit*>(ctx->scratch[3].ptr), n);
	ctx->launches += 2;
	RZB_CUDA(ctx, cudaGetLastError());
	RZB_CUDA(ctx, cudaMemcpyAsync(hits_out, ctx->scratch[3].ptr, size_t(n) * sizeof(rzb_hit), cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (stats)
	{
		unsigned long long h[4] = {};
		RZB_CUDA(ctx, cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost));
		stats->rays = n;
		stats->top_nodes = h[0]; stats->instances_entered = h[1]; stats->mesh_nodes = h[2]; stats->triangles = h[3];
	}
	return RZB_OK;
}

extern "C" int rzb_trace_any(rzb_ctx* ctx, const float* origins, const float* directions,
	const float* near_far, uint32_t n, float* mask_out)
{
	if (!ctx || !origins || !directions || !near_far || !mask_out) return fail(ctx, RZB_ERR_INVALID, "rzb_trace_any: NULL argument");
	if (!ctx->has_scene) return fail(ctx, RZB_ERR_STATE, "rzb_trace_any: no scene");
	if (n == 0) return RZB_OK;
	DeviceGuard guard(ctx->device);
	int rc;
	if ((rc = packRays(ctx, origins, directions, near_far, n))) return rc;
	if ((rc = ensureScratch(ctx, 2, size_t(n) * 16))) return rc;
	RZB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 4, ctx->stream));
	k_trace_any_rays<false><<<ctx->any_grid, kTraceBlock, 0, ctx->stream>>>(ctx->sc,
		static_cast<const float4*>(ctx->scratch[0].ptr), static_cast<const float4*>(ctx->scratch[1].ptr), n,
		static_cast<float4*>(ctx->scratch[2].ptr), ctx->d_counters + 8);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	RZB_CUDA(ctx, cudaMemcpyAsync(mask_out, ctx->scratch[2].ptr, size_t(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return RZB_OK;
}

extern "C" int rzb_generate_camera_rays(rzb_ctx* ctx, float* origins, float* directions, float* near_far)
{
	if (!ctx || !origins || !directions || !near_far) return fail(ctx, RZB_ERR_INVALID, "rzb_generate_camera_rays: NULL argument");
	if (!ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_generate_camera_rays: no camera");
	DeviceGuard guard(ctx->device);
	const uint32_t n = ctx->cam.width * ctx->cam.height;
	int rc;
	if ((rc = ensureScratch(ctx, 0, size_t(n) * 16))) return rc;
	if ((rc = ensureScratch(ctx, 1, size_t(n) * 16))) return rc;
	k_camera_rays<<<(n + 255) / 256, 256, 0, ctx->stream>>>(makeDeviceCamera(ctx->cam),
		static_cast<float4*>(ctx->scratch[0].ptr), static_cast<float4*>(ctx->scratch[1].ptr));
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	std::vector<float4> o(n), d(n);
	RZB_CUDA(ctx, cudaMemcpyAsync(o.data(), ctx->scratch[0].ptr, size_t(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaMemcpyAsync(d.data(), ctx->scratch[1].ptr, size_t(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (uint32_t i = 0; i < n; ++i)
	{
		origins[3 * size_t(i)] = o[i].x; origins[3 * size_t(i) + 1] = o[i].y; origins[3 * size_t(i) + 2] = o[i].z;
		directions[3 * size_t(i)] = d[i].x; directions[3 * size_t(i) + 1] = d[i].y; directions[3 * size_t(i) + 2] = d[i].z;
		near_far[2 * size_t(i)] = o[i].w; near_far[2 * size_t(i) + 1] = d[i].w;
	}
	return RZB_OK;
}

extern "C" int rzb_raycast(rzb_ctx* ctx, uint32_t* instance, uint32_t* material_slot)
{
	if (!ctx || !instance || !material_slot) return fail(ctx, RZB_ERR_INVALID, "rzb_raycast: NULL argument");
	if (!ctx->has_scene || !ctx->has_camera) return fail(ctx, RZB_ERR_STATE, "rzb_raycast: scene and camera must be set first");
	*instance = RZB_NO_INDEX;
	*material_slot = RZB_NO_INDEX;
	const uint32_t px = std::min(ctx->cam.raycast_pixel[0], ctx->cam.width - 1u);
	const uint32_t py = std::min(ctx->cam.raycast_pixel[1], ctx->cam.height - 1u);
	DeviceGuard guard(ctx->device);
	uint32_t* d_out = ctx->d_counters + 40;
	k_raycast<<<1, 32, 0, ctx->stream>>>(ctx->sc, makeDeviceCamera(ctx->cam), ctx->frame.depth, px, py, d_out);
	ctx->launches += 1;
	RZB_CUDA(ctx, cudaGetLastError());
	uint32_t h[2] = {RZB_NO_INDEX, RZB_NO_INDEX};
	RZB_CUDA(ctx, cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	RZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	*instance = h[0];
	*material_slot = h[1];
	return RZB_OK;
}
