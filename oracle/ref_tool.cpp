// rz_ref_tool — drives the REFERENCE'S OWN CPU implementation (compiled in place from
// /root/reference/RayZath by oracle/Makefile) for parity vectors and the CPU baseline.
//
// TEST INFRASTRUCTURE ONLY: nothing in the product path links or executes this. It is used by
// tests/, tests/tools/make_golden.py and bench.py's reference arm / cpu_baseline leg.
//
// commands
//   dumpscene <scene.json> <out.rzs>            flattened world (C-ABI arrays) + camera 0 + reference pixel-centre rays
//   trace     <scene.json> <rays.rzs> <out.rzs> closest hit per ray through CPU::Kernel (cpu_engine_kernel.cpp:254-352)
//   traceany  <scene.json> <rays.rzs> <out.rzs> shadow mask per ray through CPU::Kernel::anyIntersection (:398-481)
//   render    <scene.json> <passes> <out.rzs|-> [max_depth] [spot_samples] [direct_samples] [warmup=1]
//                                               <passes> x Engine::renderWorld(CPU), the first <warmup> of them untimed (the first call
//                                               also builds the BVHs) ; dumps float accumulator, RGBA8, depth ; prints timing JSON
//   headless  <tasks.json> [report_dir] [-r]    the reference's own Application/headless.cpp entry
//   rendercuda <scene.json> <calls> <rpp> <out.rzs|-> [max_depth] [spot_samples] [direct_samples] [warmup_calls=1]
//                                               (rz_ref_tool_cuda only) <calls> x Engine::renderWorld(CUDAGPU) with <rpp> passes each:
//                                               the reference's own CUDA engine; dumps RGBA8 + depth ; prints timing JSON
//   tracecuda <scene.json> <rays.rzs> <out.rzs> (rz_ref_tool_cuda only) closest-hit records from the reference's own CUDA
//                                               traversal (cuda_world.cuh:80-90 on the device World the engine mirrored);
//                                               same record format as `trace`
//   movecuda <scene.json> <calls> <rpp> <out.rzs> <dx> <dy> <dz> [max_depth]
//                                               (rz_ref_tool_cuda only) <calls> x renderWorld(CUDAGPU), camera moved by
//                                               (dx,dy,dz), ONE more renderWorld call: the restart blends the replaced
//                                               frame in (Camera::reproject, cuda_camera.cuh:390-426); dumps both images
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <string>
#include <thread>

// reach the CPU kernel's private traversal entry points and the accumulator (test tool only)
#define private public
#define protected public
#include "rayzath.hpp"
#include "cpu_engine.hpp"
#include "cpu_engine_core.hpp"
#include "cpu_engine_renderer.hpp"
#include "cpu_engine_kernel.hpp"
#include "cpu_render_utils.hpp"
#undef private
#undef protected

#include "../rayzath_b200/host/world_flatten.hpp"
#include "rzs_io.hpp"

#ifdef RZ_WITH_HEADLESS
#include "headless.hpp"
#endif
#ifdef RZ_WITH_CUDA_ENGINE
#include "ref_trace_cuda.h"
#endif

namespace RZ = RayZath::Engine;

static RZ::World& loadWorld(const std::string& scene_path)
{
	auto& engine = RZ::Engine::instance();
	auto& world = engine.world();
	world.loader().loadScene(scene_path);
	auto& cameras = world.container<RZ::ObjectType::Camera>();
	for (uint32_t i = 0; i < cameras.count(); ++i)
		if (cameras[i]) cameras[i]->update();
	world.update();
	return world;
}

static void writeScene(rzs::Writer& w, const rzb_host::FlatScene& s)
{
	w.add("mesh_nodes", s.mesh_nodes);
	w.add("triangles", s.triangles);
	w.add("tri_host_index", s.tri_host_index);
	w.add("meshes", s.meshes);
	w.add("instance_nodes", s.instance_nodes);
	w.add("instances", s.instances);
	w.add("instance_materials", s.instance_materials);
	w.add("materials", s.materials);
	std::vector<rzb_map> maps = s.maps;
	for (size_t i = 0; i < maps.size(); ++i)
	{
		const size_t texel = maps[i].format == RZB_MAP_R8 ? 1 : 4;
		w.add("map_pixels_" + std::to_string(i), maps[i].pixels, uint32_t(texel), uint64_t(maps[i].width) * maps[i].height);
		maps[i].pixels = nullptr;
	}
	w.add("maps", maps);
	w.add("direct_lights", s.direct_lights);
	w.add("spot_lights", s.spot_lights);
	w.addValue("world_material", s.world_material);
	w.addValue("default_material", s.default_material);
}

static int cmdDumpScene(const std::string& scene, const std::string& out_path)
{
	auto& world = loadWorld(scene);
	rzb_host::FlatScene flat;
	rzb_host::WorldFlattener(world, flat).run();
	rzs::Writer w;
	writeScene(w, flat);

	auto& cameras = world.container<RZ::ObjectType::Camera>();
	if (cameras.count() != 0 && cameras[0])
	{
		auto& cam = *cameras[0];
		w.addValue("camera", rzb_host::flattenCamera(cam));
		// the fixed primary-ray set: the reference's own pixel-centre rays (cpu_engine_kernel.cpp:180-204)
		RZ::CPU::Kernel kernel;
		kernel.setWorld(world);
		const uint32_t W = cam.width(), H = cam.height();
		std::vector<float> o(size_t(W) * H * 3), d(size_t(W) * H * 3), nf(size_t(W) * H * 2);
		for (uint32_t y = 0; y < H; ++y)
			for (uint32_t x = 0; x < W; ++x)
			{
				RZ::CPU::RangedRay ray;
				kernel.generateSimpleRay(cam, ray, Math::vec2ui32(x, y));
				const size_t i = size_t(y) * W + x;
				o[3 * i] = ray.origin.x; o[3 * i + 1] = ray.origin.y; o[3 * i + 2] = ray.origin.z;
				d[3 * i] = ray.direction.x; d[3 * i + 1] = ray.direction.y; d[3 * i + 2] = ray.direction.z;
				nf[2 * i] = ray.near_far.x; nf[2 * i + 1] = ray.near_far.y;
			}
		w.add("ray_origins", o.data(), 12, size_t(W) * H);
		w.add("ray_directions", d.data(), 12, size_t(W) * H);
		w.add("ray_near_far", nf.data(), 8, size_t(W) * H);
	}
	w.write(out_path);
	std::printf("{\"meshes\": %zu, \"triangles\": %zu, \"mesh_nodes\": %zu, \"instances\": %zu, \"instance_nodes\": %zu}\n",
		flat.meshes.size(), flat.triangles.size(), flat.mesh_nodes.size(), flat.instances.size(), flat.instance_nodes.size());
	return 0;
}

struct RaySet
{
	size_t n = 0;
	const float* o = nullptr;
	const float* d = nullptr;
	const float* nf = nullptr;
};
static RaySet raysOf(const std::map<std::string, rzs::Array>& arrays)
{
	RaySet r;
	const auto& o = arrays.at("ray_origins");
	r.n = o.count;
	r.o = o.as<float>();
	r.d = arrays.at("ray_directions").as<float>();
	r.nf = arrays.at("ray_near_far").as<float>();
	return r;
}

static int cmdTrace(const std::string& scene, const std::string& rays_path, const std::string& out_path, bool any)
{
	auto& world = loadWorld(scene);
	const auto arrays = rzs::read(rays_path);
	const RaySet rays = raysOf(arrays);
	RZ::CPU::Kernel kernel;
	kernel.setWorld(world);
	const auto& instances = world.container<RZ::ObjectType::Instance>();

	std::vector<rzb_hit> hits;
	std::vector<float> masks;
	if (any) masks.resize(rays.n * 4);
	else hits.resize(rays.n);

	const unsigned n_threads = std::max(1u, std::thread::hardware_concurrency());
	std::vector<std::thread> threads;
	const auto t0 = std::chrono::steady_clock::now();
	for (unsigned t = 0; t < n_threads; ++t)
		threads.emplace_back([&, t]() {
			for (size_t i = t; i < rays.n; i += n_threads)
			{
				RZ::CPU::RangedRay ray;
				ray.origin = Math::vec3f32(rays.o[3 * i], rays.o[3 * i + 1], rays.o[3 * i + 2]);
				ray.direction = Math::vec3f32(rays.d[3 * i], rays.d[3 * i + 1], rays.d[3 * i + 2]);
				ray.near_far = Math::vec2f32(rays.nf[2 * i], rays.nf[2 * i + 1]);
				if (any)
				{
					const Graphics::ColorF m = kernel.anyIntersection(ray);
					masks[4 * i] = m.red; masks[4 * i + 1] = m.green; masks[4 * i + 2] = m.blue; masks[4 * i + 3] = m.alpha;
					continue;
				}
				rzb_hit h{};
				h.instance = RZB_NO_INDEX; h.triangle = RZB_NO_INDEX;
				// Kernel::closestIntersection(RangedRay&, SurfaceProperties&) without the surface analysis
				// (cpu_engine_kernel.cpp:279-289)
				RZ::CPU::TraversalResult traversal;
				if (!instances.empty() && instances.root().boundingBox().rayIntersection(ray))
					kernel.traverseWorld(instances.root(), ray, traversal);
				if (traversal.closest_instance)
				{
					h.instance = traversal.instance_idx;
					const auto& mesh = *traversal.closest_instance->mesh();
					h.triangle = uint32_t(traversal.closest_triangle - &mesh.triangles()[0]);
					h.b1 = traversal.barycenter.x; h.b2 = traversal.barycenter.y;
					h.external = traversal.external ? 1u : 0u;
				}
				h.t = ray.near_far.y;
				hits[i] = h;
			}
		});
	for (auto& th : threads) th.join();
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	rzs::Writer w;
	if (any) w.add("masks", masks.data(), 16, rays.n);
	else w.add("hits", hits);
	w.write(out_path);
	std::printf("{\"rays\": %zu, \"seconds\": %.6f, \"threads\": %u}\n", rays.n, secs, n_threads);
	return 0;
}

static int cmdRender(int argc, char** argv)
{
	const std::string scene = argv[2];
	const uint32_t passes = uint32_t(std::atoi(argv[3]));
	const std::string out_path = argv[4];
	auto& engine = RZ::Engine::instance();
	auto& world = engine.world();
	const auto t_load0 = std::chrono::steady_clock::now();
	world.loader().loadScene(scene);
	if (argc > 5) engine.renderConfig().tracing().maxDepth(uint8_t(std::atoi(argv[5])));
	if (argc > 6) engine.renderConfig().lightSampling().spotLight(uint8_t(std::atoi(argv[6])));
	if (argc > 7) engine.renderConfig().lightSampling().directLight(uint8_t(std::atoi(argv[7])));
	const uint32_t warmup = std::min(passes, argc > 8 ? uint32_t(std::atoi(argv[8])) : 1u);
	engine.renderEngine(RZ::Engine::RenderEngine::CPU);

	// first call: world.update() (BVH build) + first pass; timed separately as in headless.cpp:209
	for (uint32_t p = 0; p < warmup; ++p)
		engine.renderWorld(RZ::Engine::RenderEngine::CPU, true, true);
	const double load_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load0).count();
	const auto t0 = std::chrono::steady_clock::now();
	for (uint32_t p = warmup; p < passes; ++p)
		engine.renderWorld(RZ::Engine::RenderEngine::CPU, true, true);
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	auto& cameras = world.container<RZ::ObjectType::Camera>();
	uint64_t rays = 0;
	rzs::Writer w;
	for (uint32_t i = 0; i < cameras.count(); ++i)
	{
		if (!cameras[i]) continue;
		rays += cameras[i]->rayCount();
		if (i != 0) continue;
		auto& cam = *cameras[i];
		auto& ctx = engine.m_cpu_engine->m_engine_core.m_renderer.m_contexts[cameras[i]];
		w.add("accum", ctx.m_image.GetMapAddress(), 16, uint64_t(cam.width()) * cam.height());
		w.add("rgba8", cam.imageBuffer().GetMapAddress(), 4, uint64_t(cam.width()) * cam.height());
		w.add("depth", cam.depthBuffer().GetMapAddress(), 4, uint64_t(cam.width()) * cam.height());
		const uint32_t res[2] = {cam.width(), cam.height()};
		w.add("resolution", res, 4, 2);
	}
	if (out_path != "-") w.write(out_path);
	const uint64_t timed_rays = passes ? rays / passes * (passes - warmup) : 0;
	std::printf("{\"passes\": %u, \"rays\": %llu, \"timed_passes\": %u, \"timed_rays\": %llu, \"seconds\": %.6f, "
		"\"first_call_seconds\": %.6f, \"threads\": %u}\n",
		passes, (unsigned long long)rays, passes - warmup, (unsigned long long)timed_rays, secs, load_secs,
		std::thread::hardware_concurrency());
	return 0;
}

#ifdef RZ_WITH_CUDA_ENGINE
static int cmdRenderCuda(int argc, char** argv)
{
	const std::string scene = argv[2];
	const uint32_t calls = uint32_t(std::atoi(argv[3]));
	const uint32_t rpp = uint32_t(std::atoi(argv[4]));
	const std::string out_path = argv[5];
	auto& engine = RZ::Engine::instance();
	auto& world = engine.world();
	world.loader().loadScene(scene);
	if (argc > 6) engine.renderConfig().tracing().maxDepth(uint8_t(std::atoi(argv[6])));
	if (argc > 7) engine.renderConfig().lightSampling().spotLight(uint8_t(std::atoi(argv[7])));
	if (argc > 8) engine.renderConfig().lightSampling().directLight(uint8_t(std::atoi(argv[8])));
	const uint32_t warmup = std::min(calls, argc > 9 ? uint32_t(std::atoi(argv[9])) : 1u);
	engine.renderConfig().tracing().rpp(rpp);
	if (engine.renderEngine() != RZ::Engine::RenderEngine::CUDAGPU)
	{
		std::fprintf(stderr, "rz_ref_tool: the reference CUDA engine failed to initialise (no GPU?)\n");
		return 3;
	}
	auto& cameras = world.container<RZ::ObjectType::Camera>();
	auto rayCount = [&]() {
		uint64_t rays = 0;
		for (uint32_t i = 0; i < cameras.count(); ++i) if (cameras[i]) rays += cameras[i]->rayCount();
		return rays;
	};
	for (uint32_t c = 0; c < warmup; ++c) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
	const uint64_t rays0 = rayCount();
	const auto t0 = std::chrono::steady_clock::now();
	for (uint32_t c = warmup; c < calls; ++c) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	const uint64_t rays1 = rayCount();
	rzs::Writer w;
	if (cameras.count() && cameras[0])
	{
		auto& cam = *cameras[0];
		w.add("rgba8", cam.imageBuffer().GetMapAddress(), 4, uint64_t(cam.width()) * cam.height());
		w.add("depth", cam.depthBuffer().GetMapAddress(), 4, uint64_t(cam.width()) * cam.height());
		const uint32_t res[2] = {cam.width(), cam.height()};
		w.add("resolution", res, 4, 2);
	}
	if (out_path != "-") w.write(out_path);
	std::printf("{\"calls\": %u, \"rpp\": %u, \"timed_calls\": %u, \"timed_rays\": %llu, \"rays\": %llu, \"seconds\": %.6f}\n",
		calls, rpp, calls - warmup, (unsigned long long)(rays1 - rays0), (unsigned long long)rays1, secs);
	// the engine singleton is destroyed after the CUDA runtime has shut down (the reference throws from its destructor
	// then: "driver shutting down"); results are complete, so leave without running static destructors
	std::fflush(stdout);
	std::_Exit(0);
}
#endif

#ifdef RZ_WITH_CUDA_ENGINE
static int cmdTraceCuda(const std::string& scene, const std::string& rays_path, const std::string& out_path)
{
	auto& engine = RZ::Engine::instance();
	auto& world = engine.world();
	world.loader().loadScene(scene);
	engine.renderConfig().tracing().rpp(1);
	if (engine.renderEngine() != RZ::Engine::RenderEngine::CUDAGPU)
	{
		std::fprintf(stderr, "rz_ref_tool: the reference CUDA engine failed to initialise (no GPU?)\n");
		return 3;
	}
	// one call mirrors the world to the device exactly as the reference does (World::reconstructAll)
	engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
	// BVH-order triangle index of the device mesh -> host triangle index: Mesh::reconstruct emits leaf by leaf in the
	// order world_flatten.hpp restates (cuda_instance.cu:101-220)
	rzb_host::FlatScene flat;
	rzb_host::WorldFlattener(world, flat).run();
	std::vector<uint32_t> mesh_of_instance(world.container<RZ::ObjectType::Instance>().count(), RZB_NO_INDEX);
	for (const auto& in : flat.instances)
		if (in.host_index < mesh_of_instance.size()) mesh_of_instance[in.host_index] = in.mesh;

	const auto arrays = rzs::read(rays_path);
	const RaySet rays = raysOf(arrays);
	std::vector<RefCudaHit> raw(rays.n);
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = refCudaTrace(engine.m_cuda_engine.get(), rays.o, rays.d, rays.nf, uint32_t(rays.n), raw.data());
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (rc) return 4;
	std::vector<rzb_hit> hits(rays.n);
	for (size_t i = 0; i < rays.n; ++i)
	{
		rzb_hit h{};
		h.instance = RZB_NO_INDEX; h.triangle = RZB_NO_INDEX;
		h.t = raw[i].t;
		if (raw[i].instance != 0xFFFFFFFFu)
		{
			h.instance = raw[i].instance;
			const uint32_t m = raw[i].instance < mesh_of_instance.size() ? mesh_of_instance[raw[i].instance] : RZB_NO_INDEX;
			if (m != RZB_NO_INDEX && raw[i].triangle_bvh_order < flat.meshes[m].tri_count)
				h.triangle = flat.tri_host_index[flat.meshes[m].tri_offset + raw[i].triangle_bvh_order];
			h.b1 = raw[i].b1; h.b2 = raw[i].b2; h.external = raw[i].external;
		}
		hits[i] = h;
	}
	rzs::Writer w;
	w.add("hits", hits);
	w.write(out_path);
	std::printf("{\"rays\": %zu, \"seconds\": %.6f}\n", rays.n, secs);
	std::fflush(stdout);
	std::_Exit(0);
}

static int cmdMoveCuda(int argc, char** argv)
{
	const std::string scene = argv[2];
	const uint32_t calls = uint32_t(std::atoi(argv[3]));
	const uint32_t rpp = uint32_t(std::atoi(argv[4]));
	const std::string out_path = argv[5];
	const Math::vec3f delta(float(std::atof(argv[6])), float(std::atof(argv[7])), float(std::atof(argv[8])));
	auto& engine = RZ::Engine::instance();
	auto& world = engine.world();
	world.loader().loadScene(scene);
	if (argc > 9) engine.renderConfig().tracing().maxDepth(uint8_t(std::atoi(argv[9])));
	engine.renderConfig().tracing().rpp(rpp);
	if (engine.renderEngine() != RZ::Engine::RenderEngine::CUDAGPU) return 3;
	auto& cameras = world.container<RZ::ObjectType::Camera>();
	if (!cameras.count() || !cameras[0]) return 5;
	auto& cam = *cameras[0];
	const uint64_t n = uint64_t(cam.width()) * cam.height();
	rzs::Writer w;
	for (uint32_t c = 0; c < calls; ++c) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
	std::vector<uint8_t> before(n * 4);
	std::memcpy(before.data(), cam.imageBuffer().GetMapAddress(), n * 4);
	w.add("rgba8_before", before.data(), 4, n);
	cam.position(cam.position() + delta);
	engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
	w.add("rgba8_after", cam.imageBuffer().GetMapAddress(), 4, n);
	const uint32_t res[2] = {cam.width(), cam.height()};
	w.add("resolution", res, 4, 2);
	w.write(out_path);
	std::printf("{\"calls\": %u, \"rpp\": %u, \"rays_after_move\": %llu}\n", calls, rpp, (unsigned long long)cam.rayCount());
	std::fflush(stdout);
	std::_Exit(0);
}
#endif

int main(int argc, char** argv)
{
	try
	{
		const std::string cmd = argc > 1 ? argv[1] : "";
		if (cmd == "dumpscene" && argc == 4) return cmdDumpScene(argv[2], argv[3]);
		if (cmd == "trace" && argc == 5) return cmdTrace(argv[2], argv[3], argv[4], false);
		if (cmd == "traceany" && argc == 5) return cmdTrace(argv[2], argv[3], argv[4], true);
		if (cmd == "render" && argc >= 5) return cmdRender(argc, argv);
#ifdef RZ_WITH_CUDA_ENGINE
		if (cmd == "rendercuda" && argc >= 6) return cmdRenderCuda(argc, argv);
		if (cmd == "tracecuda" && argc == 5) return cmdTraceCuda(argv[2], argv[3], argv[4]);
		if (cmd == "movecuda" && argc >= 9) return cmdMoveCuda(argc, argv);
#endif
#ifdef RZ_WITH_HEADLESS
		if (cmd == "headless" && argc >= 3)
		{
			std::filesystem::path report = argc > 3 && std::string(argv[3]) != "-r" ? argv[3] : "";
			bool save = false;
			for (int i = 3; i < argc; ++i) save |= std::string(argv[i]) == "-r";
			return RayZath::Headless::Headless::instance().run(argv[2], report, save);
		}
#endif
		std::fprintf(stderr, "usage: rz_ref_tool dumpscene|trace|traceany|render|headless ... (see oracle/ref_tool.cpp)\n");
		return 2;
	}
	catch (const std::exception& e)
	{
		std::fprintf(stderr, "rz_ref_tool: %s\n", e.what());
		return 1;
	}
}
