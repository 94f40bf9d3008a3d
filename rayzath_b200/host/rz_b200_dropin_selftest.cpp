// Self-test driver of the drop-in (tests/test_gpu_dropin.py): the reference's host API with world edits BETWEEN
// frames, which the headless runner never does -- exercises the dirty-flag protocol of cuda_engine_b200.cpp
// (incremental upload when only instances / materials changed, accumulation restart, temporal reprojection history,
// sync == true and the pipelined sync == false).
//   rz_b200_dropin_selftest <scene.json> <out.raw>     writes three RGBA8 frames back to back:
//     frame 0: 4 renderWorld calls on the loaded scene (sync = true)
//     frame 1: every instance moved by (0.25, 0.1, 0), first material recoloured, 4 calls (sync = true)
//     frame 2: 4 more calls with sync = false, then one sync = true call
//     frame 3 (only with a third argument N): N more instances of the first mesh are created -- no mesh edit, so the
//              upload is incremental until the instance tree outgrows the room reserved for it (1024 nodes), where the
//              engine must fall back to a full upload instead of failing -- then 4 calls (sync = true)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <fstream>
#include <iostream>
#include <vector>

#include "rayzath.hpp"

namespace RZ = RayZath::Engine;

static void append(std::ofstream& out, RZ::Camera& cam)
{
	const auto& img = cam.imageBuffer();
	out.write(reinterpret_cast<const char*>(img.GetMapAddress()), std::streamsize(img.GetWidth()) * img.GetHeight() * 4);
}

int main(int argc, char* argv[])
{
	if (argc < 3) { std::fprintf(stderr, "usage: %s scene.json out.raw\n", argv[0]); return 2; }
	try
	{
		auto& engine = RZ::Engine::instance();
		auto& world = engine.world();
		world.loader().loadScene(argv[1]);
		engine.renderConfig().tracing().rpp(8);
		auto& cameras = world.container<RZ::ObjectType::Camera>();
		if (cameras.count() == 0 || !cameras[0]) throw std::runtime_error("scene without a camera");
		std::ofstream out(argv[2], std::ios::binary);

		for (int i = 0; i < 4; ++i) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
		append(out, *cameras[0]);

		auto& instances = world.container<RZ::ObjectType::Instance>();
		for (uint32_t i = 0; i < instances.count(); ++i)
			if (instances[i]) instances[i]->position(instances[i]->transformation().position() + Math::vec3f(0.25f, 0.1f, 0.0f));
		auto& materials = world.container<RZ::ObjectType::Material>();
		if (materials.count() && materials[0]) materials[0]->color(Graphics::Color(20, 200, 40, 255));
		for (int i = 0; i < 4; ++i) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
		append(out, *cameras[0]);

		for (int i = 0; i < 4; ++i) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, false);
		engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
		append(out, *cameras[0]);
		const unsigned long long rays_frame2 = cameras[0]->rayCount();
		uint32_t grown = 0;
		if (argc > 3)
		{
			const int extra = std::atoi(argv[3]);
			RayZath::Engine::Handle<RZ::Mesh> mesh;
			RayZath::Engine::Handle<RZ::Material> mat;
			for (uint32_t i = 0; i < instances.count() && !mesh; ++i)
				if (instances[i] && instances[i]->mesh()) { mesh = instances[i]->mesh(); mat = instances[i]->material(0); }
			for (int i = 0; i < extra; ++i)
			{
				const float x = float(i % 40) * 0.3f - 6.0f, z = float(i / 40) * 0.3f - 1.0f;
				instances.create(RZ::ConStruct<RZ::Instance>("grown " + std::to_string(i), Math::vec3f(x, 2.5f, z),
					Math::vec3f(0.0f, 0.0f, 0.0f), Math::vec3f(0.1f, 0.1f, 0.1f), mesh, mat));
			}
			for (int i = 0; i < 4; ++i) engine.renderWorld(RZ::Engine::RenderEngine::CUDAGPU, true, true);
			append(out, *cameras[0]);
			grown = instances.count();
		}
		std::printf("{\"width\": %u, \"height\": %u, \"rays\": %llu, \"instances_after_growth\": %u, \"rays_after_growth\": %llu}\n",
			cameras[0]->width(), cameras[0]->height(), rays_frame2, grown, (unsigned long long)cameras[0]->rayCount());
		return 0;
	}
	catch (const std::exception& e)
	{
		std::fprintf(stderr, "selftest: %s\n", e.what());
		return 1;
	}
}
