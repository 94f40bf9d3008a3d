// Drop-in implementation of RayZath::Cuda::Engine on top of the B200 render path's C ABI (include/rzb200.h).
//
// The reference's facade keeps `std::unique_ptr<RayZath::Cuda::Engine> m_cuda_engine` (rayzath.hpp:31) and calls
//   m_cuda_engine->renderWorld(*m_world, m_render_config, block, sync)      rayzath.cpp:64-90
//   m_cuda_engine->timingsString()                                          rayzath.cpp:96-113
// This file defines that class (declaration unchanged: /root/reference/RayZath/cuda_engine.cuh:21-40) and the
// `RayZath::Cuda::EngineCore` the reference's Camera befriends (camera.hpp:121), so linking it INSTEAD of the
// reference's cuda_*.cu files replaces the whole CUDA engine while rayzath.hpp, World, RenderConfig, json_loader
// and the headless entry stay byte-for-byte the reference's. What it replaces:
//   EngineCore::renderWorld / CopyRenderToHost    cuda_engine_core.cu:32-240
//   Renderer::renderFunction                      cuda_engine_renderer.cu:73-262
//   World/Mesh/Instance/...::reconstruct          cuda_world.cu, cuda_instance.cu, cuda_bvh.cuh (-> world_flatten.hpp)
//
// Behaviour kept: dirty-flag protocol (world modified -> re-mirror, camera modified -> restart accumulation,
// flags cleared afterwards), `rpp` passes per call, results written to Camera::m_image_buffer / m_depth_buffer /
// m_ray_count / m_raycasted_*, errors as RayZath::Cuda::Exception, CUDA init failure -> the facade falls back to
// the CPU engine (rayzath.cpp:21-28); sync == false pipelines one frame as the reference does (the caller gets the
// previous frame while this one renders; single device only). Behaviour changed: the per-call random seeds of
// cuda_kernel_data.cu:10-18 become one seed per engine (RZB200_SEED or std::random_device) + the pass counter.
// RZB200_BVH=sah replaces the host World's triangle trees by this repo's SAH trees (RZB_SCENE_OWN_TREES; hit records
// equal except exact-distance ties). Multi-GPU: RZB200_DEVICES="0,1,..." renders one disjoint sample stream per listed device and sums the
// accumulators in the fused peer resolve (rzb_resolve_peers) on the first device.
#include "cuda_engine.cuh"
#include "cuda_exception.hpp"

#include "world_flatten.hpp"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <random>
#include <sstream>

namespace RayZath::Cuda
{
	namespace RZ = RayZath::Engine;

	class EngineCore;
	static EngineCore* g_core = nullptr; // the live engine, for the headless runner's hooks (rzb200_engine_*)

	class EngineCore
	{
		struct CameraState
		{
			std::vector<rzb_ctx*> ctxs; // one per device
			uint32_t width = 0, height = 0;
			bool scene_current = false;
			uint64_t geometry_version = 0; // m_geometry_version of the last full upload into these contexts
			std::vector<uint8_t> rgba;
			std::vector<float> depth;
			// sync == false: one frame in flight (pinned double buffer), the caller gets the previous one
			void* pin_rgba[2] = {nullptr, nullptr};
			void* pin_depth[2] = {nullptr, nullptr};
			size_t pin_pixels = 0;
			int pending = -1;          // slot whose asynchronous resolve has not been handed back yet
			uint64_t pending_rays = 0;
			uint64_t frame_no = 0;
		};
		std::mutex m_mtx;
		std::vector<int> m_devices;
		std::map<uint32_t, CameraState> m_cameras; // by camera container index
		uint64_t m_seed = 0;
		uint64_t m_scene_version = 0;
		uint64_t m_geometry_version = 0; // bumped when the Mesh container changed: everything else updates incrementally
		bool m_full_upload = false; // RZB200_FULL_UPLOAD=1: re-upload the geometry on every world change (test aid)
		bool m_own_trees = false; // RZB200_BVH=sah | sah4: this repo's SAH triangle trees instead of the host World's
		bool m_wide_trees = false; // RZB200_BVH=sah4: ... collapsed to 4-ary trees by rzb_set_scene (RZB_SCENE_WIDE_TREES)
		rzb_host::FlatScene m_flat;
		std::string m_timings;

		[[noreturn]] static void fail(rzb_ctx* ctx, const char* what)
		{
			throw Exception(std::string(what) + ": " + rzb_last_error(ctx));
		}
		static void check(rzb_ctx* ctx, int rc, const char* what)
		{
			if (rc != RZB_OK) fail(ctx, what);
		}

	public:
		EngineCore()
		{
			if (const char* env = std::getenv("RZB200_DEVICES"))
			{
				std::stringstream ss(env);
				for (std::string tok; std::getline(ss, tok, ',');)
					if (!tok.empty()) m_devices.push_back(std::atoi(tok.c_str()));
			}
			if (m_devices.empty()) m_devices.push_back(0);
			if (const char* env = std::getenv("RZB200_BVH"))
			{
				m_own_trees = std::string(env) == "sah" || std::string(env) == "sah4";
				m_wide_trees = std::string(env) == "sah4";
			}
			if (const char* env = std::getenv("RZB200_FULL_UPLOAD")) m_full_upload = std::string(env) == "1";
			if (const char* env = std::getenv("RZB200_SEED")) m_seed = std::strtoull(env, nullptr, 0);
			else
			{
				std::random_device rd;
				m_seed = (uint64_t(rd()) << 32) | rd();
			}
			if (rzb_abi_version() != int(RZB_ABI_VERSION))
				throw Exception("librzb200.so has ABI version " + std::to_string(rzb_abi_version()) + ", this engine was built for " + std::to_string(RZB_ABI_VERSION));
			// probe the first device now: a missing GPU must surface from the constructor so that
			// RayZath::Engine::Engine falls back to the CPU engine
			rzb_ctx* probe = nullptr;
			if (rzb_create(m_devices[0], &probe) != RZB_OK) fail(nullptr, "B200 render path unavailable");
			rzb_destroy(probe);
			g_core = this;
		}
		// ---- hooks of the headless task format (task keys "seed", "devices", "spp", "accumulator")
		void configure(const bool has_seed, const uint64_t seed, const char* devices)
		{
			std::lock_guard<std::mutex> lg(m_mtx);
			if (has_seed) m_seed = seed;
			if (devices && *devices)
			{
				std::vector<int> list;
				std::stringstream ss(devices);
				for (std::string tok; std::getline(ss, tok, ',');)
					if (!tok.empty()) list.push_back(std::atoi(tok.c_str()));
				if (!list.empty() && list != m_devices)
				{
					dropContexts();
					m_devices = list;
				}
			}
			// a new task starts from a clean accumulation with the new seed
			for (auto& [idx, cam] : m_cameras) cam.scene_current = false;
		}
		double meanSpp(const uint32_t camera)
		{
			std::lock_guard<std::mutex> lg(m_mtx);
			auto it = m_cameras.find(camera);
			if (it == m_cameras.end()) return 0.0;
			double sum = 0.0;
			for (rzb_ctx* c : it->second.ctxs)
			{
				double m = 0.0;
				check(c, rzb_mean_samples(c, &m), "rzb_mean_samples");
				sum += m; // sample streams of several devices add up
			}
			return sum;
		}
		bool readAccum(const uint32_t camera, float* out, const size_t floats)
		{
			std::lock_guard<std::mutex> lg(m_mtx);
			auto it = m_cameras.find(camera);
			if (it == m_cameras.end() || it->second.ctxs.empty()) return false;
			const size_t n = size_t(it->second.width) * it->second.height * 4;
			if (floats < n) return false;
			std::vector<float> tmp;
			for (size_t d = 0; d < it->second.ctxs.size(); ++d)
			{
				if (d == 0) { check(it->second.ctxs[0], rzb_read_accum(it->second.ctxs[0], out), "rzb_read_accum"); continue; }
				tmp.resize(n);
				check(it->second.ctxs[d], rzb_read_accum(it->second.ctxs[d], tmp.data()), "rzb_read_accum");
				for (size_t i = 0; i < n; ++i) out[i] += tmp[i];
			}
			return true;
		}
		void dropContexts()
		{
			for (auto& [idx, cam] : m_cameras)
			{
				for (rzb_ctx* c : cam.ctxs) rzb_destroy(c);
				for (int k = 0; k < 2; ++k) { rzb_host_free(cam.pin_rgba[k]); rzb_host_free(cam.pin_depth[k]); }
			}
			m_cameras.clear();
		}
		~EngineCore()
		{
			if (g_core == this) g_core = nullptr;
			// RZB200_VERBOSE: leave a trace on stderr that the B200 path (not the CPU fallback) rendered
			if (std::getenv("RZB200_VERBOSE") && !m_timings.empty())
				std::fprintf(stderr, "[rzb200] %zu device(s), seed %llu\n%s", m_devices.size(), (unsigned long long)m_seed, m_timings.c_str());
			for (auto& [idx, cam] : m_cameras)
			{
				for (rzb_ctx* c : cam.ctxs) rzb_destroy(c); // synchronises the stream: nothing is in flight afterwards
				for (int k = 0; k < 2; ++k) { rzb_host_free(cam.pin_rgba[k]); rzb_host_free(cam.pin_depth[k]); }
			}
		}

		void renderWorld(RZ::World& hWorld, const RZ::RenderConfig& config, const bool sync)
		{
			std::lock_guard<std::mutex> lg(m_mtx);

			// dirty flags, exactly as EngineCore::renderWorld reads them (cuda_engine_core.cu:48-60)
			const bool world_update = hWorld.stateRegister().IsModified();
			auto& hCameras = hWorld.container<RZ::ObjectType::Camera>();
			if (world_update)
				for (uint32_t i = 0; i < hCameras.count(); i++)
					if (hCameras[i]) hCameras[i]->stateRegister().RequestUpdate();
			hWorld.update();
			hCameras.update();

			if (world_update || m_scene_version == 0)
			{
				// incremental by dirty flags: triangles and mesh trees are re-flattened and re-uploaded only when the Mesh
				// container changed (the reference re-mirrors per container too, cuda_world.cu:40-76)
				const bool geometry_dirty = m_scene_version == 0 || m_full_upload || containerModified<RZ::ObjectType::Mesh>(hWorld);
				rzb_host::WorldFlattener(hWorld, m_flat, m_own_trees).run(!geometry_dirty);
				if (m_wide_trees && (m_flat.scene_flags & RZB_SCENE_OWN_TREES)) m_flat.scene_flags |= RZB_SCENE_WIDE_TREES;
				if (geometry_dirty) ++m_geometry_version;
				++m_scene_version;
				for (auto& [idx, cam] : m_cameras) cam.scene_current = false;
				clearFlags(hWorld);
			}
			const rzb_scene scene = m_flat.view();
			rzb_scene scene_update = scene; // same arrays, geometry kept on the device
			scene_update.flags |= RZB_SCENE_KEEP_GEOMETRY;

			for (uint32_t ci = 0; ci < hCameras.count(); ++ci)
			{
				const auto& hCamera = hCameras[ci];
				if (!hCamera || !hCamera->enabled()) continue;
				CameraState& cs = m_cameras[ci];
				if (cs.ctxs.empty())
					for (int dev : m_devices)
					{
						rzb_ctx* c = nullptr;
						if (rzb_create(dev, &c) != RZB_OK) fail(nullptr, "rzb_create");
						cs.ctxs.push_back(c);
					}
				const bool camera_update = hCamera->stateRegister().IsModified() || !cs.scene_current ||
					cs.width != hCamera->width() || cs.height != hCamera->height();
				const rzb_camera cam = rzb_host::flattenCamera(*hCamera);
				for (size_t d = 0; d < cs.ctxs.size(); ++d)
				{
					rzb_ctx* c = cs.ctxs[d];
					if (!cs.scene_current)
					{
						// geometry unchanged: send instances / materials / lights only. When the instance tree has outgrown the
						// room the last full upload reserved for it (RZB_ERR_STATE), send the whole scene, which re-reserves.
						int rc = cs.geometry_version == m_geometry_version ? rzb_set_scene(c, &scene_update) : RZB_ERR_STATE;
						if (rc == RZB_ERR_STATE) rc = rzb_set_scene(c, &scene);
						check(c, rc, "rzb_set_scene");
					}
					// one disjoint sample stream per device
					rzb_config cfg = rzb_host::flattenConfig(config, m_seed + 0x9E3779B97F4A7C15ull * (uint64_t(ci) * 64 + d));
					cfg.flags |= RZB_FLAG_TEMPORAL_REPROJECTION; // as the reference: every restart blends the replaced frame in
					check(c, rzb_set_config(c, &cfg), "rzb_set_config");
					if (camera_update)
					{
						// camera moved / world changed: restart accumulation (passReset + generateCameraRay + renderFirstPass,
						// cuda_engine_renderer.cu:87-160)
						check(c, rzb_set_camera(c, &cam), "rzb_set_camera");
						check(c, rzb_reset(c), "rzb_reset");
					}
					check(c, rzb_render(c, std::max<uint32_t>(config.tracing().rpp(), 1u)), "rzb_render");
				}
				cs.scene_current = true;
				cs.geometry_version = m_geometry_version;
				cs.width = hCamera->width();
				cs.height = hCamera->height();
				hCamera->stateRegister().MakeUnmodified();

				// copy-back (CopyRenderToHost, cuda_engine_core.cu:163-240)
				const size_t n = size_t(cs.width) * cs.height;
				uint64_t rays = 0;
				rzb_ctx* root = cs.ctxs[0];
				uint32_t inst = RZB_NO_INDEX, slot = RZB_NO_INDEX;
				const uint8_t* rgba_src = nullptr;
				const float* depth_src = nullptr;
				if (sync || cs.ctxs.size() > 1)
				{
					// synchronous: finish and hand back THIS frame (also drains a frame still in flight)
					cs.pending = -1;
					cs.rgba.resize(n * 4);
					cs.depth.resize(n);
					check(root, rzb_resolve_peers(root, cs.ctxs.data() + 1, uint32_t(cs.ctxs.size() - 1), cs.rgba.data(),
						cs.depth.data(), &rays), "rzb_resolve_peers");
					check(root, rzb_raycast(root, &inst, &slot), "rzb_raycast");
					rgba_src = cs.rgba.data();
					depth_src = cs.depth.data();
				}
				else
				{
					// pipelined, as the reference with sync == false (cuda_engine_core.cu:111-121): this frame's tone map,
					// copies and pick are enqueued behind its passes; the caller gets the previous frame, whose copies
					// finished while the host was busy. The very first frame has no predecessor and is waited for.
					if (cs.pin_pixels != n)
					{
						if (cs.pending >= 0) check(root, rzb_resolve_wait(root, uint32_t(cs.pending), nullptr, nullptr), "rzb_resolve_wait");
						cs.pending = -1;
						for (int k = 0; k < 2; ++k)
						{
							rzb_host_free(cs.pin_rgba[k]); rzb_host_free(cs.pin_depth[k]);
							cs.pin_rgba[k] = cs.pin_depth[k] = nullptr;
							if (rzb_host_alloc(n * 4, &cs.pin_rgba[k]) != RZB_OK || rzb_host_alloc(n * 4, &cs.pin_depth[k]) != RZB_OK)
								fail(root, "rzb_host_alloc");
						}
						cs.pin_pixels = n;
					}
					const int k = int(cs.frame_no & 1u);
					uint64_t rays_k = 0;
					check(root, rzb_resolve_async(root, uint32_t(k), static_cast<uint8_t*>(cs.pin_rgba[k]),
						static_cast<float*>(cs.pin_depth[k]), &rays_k), "rzb_resolve_async");
					const int give = cs.pending >= 0 ? cs.pending : k;
					check(root, rzb_resolve_wait(root, uint32_t(give), &inst, &slot), "rzb_resolve_wait");
					rays = cs.pending >= 0 ? cs.pending_rays : rays_k;
					rgba_src = static_cast<const uint8_t*>(cs.pin_rgba[give]);
					depth_src = static_cast<const float*>(cs.pin_depth[give]);
					cs.pending = k;
					cs.pending_rays = rays_k;
					++cs.frame_no;
				}
				hCamera->m_ray_count = rays;
				hCamera->m_image_buffer.CopyFromMemory(rgba_src, n * 4, 0, 0);
				hCamera->m_depth_buffer.CopyFromMemory(depth_src, n * sizeof(float), 0, 0);

				auto& hInstances = hWorld.container<RZ::ObjectType::Instance>();
				if (inst < hInstances.count() && hInstances[inst])
				{
					hCamera->m_raycasted_instance = hInstances[inst];
					if (slot < RZ::Instance::materialCapacity()) hCamera->m_raycasted_material = hInstances[inst]->material(slot);
					else hCamera->m_raycasted_material.release();
				}
				else
				{
					hCamera->m_raycasted_instance.release();
					hCamera->m_raycasted_material.release();
				}

				char buf[1024];
				if (rzb_timings(root, buf, sizeof(buf)) == RZB_OK) m_timings = buf;
			}
			hWorld.stateRegister().MakeUnmodified();
		}
		const std::string& timings() const { return m_timings; }

	private:
		// every reconstruct() of the reference ends with MakeUnmodified() on what it mirrored (e.g. cuda_instance.cu:225,
		// 281, cuda_world.cu:69-76); the GUI and the CPU engine read the same flags
		template <RZ::ObjectType T>
		static bool containerModified(RZ::World& w)
		{
			auto& c = w.container<T>();
			if (c.stateRegister().IsModified()) return true;
			for (uint32_t i = 0; i < c.count(); ++i)
				if (c[i] && c[i]->stateRegister().IsModified()) return true;
			return false;
		}
		template <RZ::ObjectType T>
		static void clearContainer(RZ::World& w)
		{
			auto& c = w.container<T>();
			for (uint32_t i = 0; i < c.count(); ++i)
				if (c[i]) c[i]->stateRegister().MakeUnmodified();
			c.stateRegister().MakeUnmodified();
		}
		static void clearFlags(RZ::World& w)
		{
			clearContainer<RZ::ObjectType::Texture>(w);
			clearContainer<RZ::ObjectType::NormalMap>(w);
			clearContainer<RZ::ObjectType::MetalnessMap>(w);
			clearContainer<RZ::ObjectType::RoughnessMap>(w);
			clearContainer<RZ::ObjectType::EmissionMap>(w);
			clearContainer<RZ::ObjectType::Material>(w);
			clearContainer<RZ::ObjectType::Mesh>(w);
			clearContainer<RZ::ObjectType::SpotLight>(w);
			clearContainer<RZ::ObjectType::DirectLight>(w);
			clearContainer<RZ::ObjectType::Instance>(w);
			w.material().stateRegister().MakeUnmodified();
			w.defaultMaterial().stateRegister().MakeUnmodified();
		}
	};

	Engine::Engine() : m_engine_core(std::make_unique<EngineCore>()) {}
}

// Hooks for the headless runner's extra task keys (linux_port/patch_headless_cpp.py declares them weak, so the same
// generated headless.cpp also links against the reference's own engines, where the keys are ignored).
extern "C" void rzb200_engine_configure(int has_seed, unsigned long long seed, const char* devices)
{
	if (RayZath::Cuda::g_core) RayZath::Cuda::g_core->configure(has_seed != 0, seed, devices);
}
extern "C" double rzb200_engine_mean_spp(unsigned camera)
{
	return RayZath::Cuda::g_core ? RayZath::Cuda::g_core->meanSpp(camera) : 0.0;
}
extern "C" int rzb200_engine_read_accum(unsigned camera, float* rgba_f32, size_t floats)
{
	return RayZath::Cuda::g_core && RayZath::Cuda::g_core->readAccum(camera, rgba_f32, floats) ? 0 : 1;
}

namespace RayZath::Cuda
{
	Engine::~Engine() = default;

	void Engine::renderWorld(RayZath::Engine::World& hWorld, const RayZath::Engine::RenderConfig& render_config,
		const bool /*block*/, const bool sync)
	{
		m_engine_core->renderWorld(hWorld, render_config, sync);
		m_timing_string = m_engine_core->timings();
	}
	std::string Engine::timingsString() { return m_timing_string; }
}
