// TEST INFRASTRUCTURE ONLY (oracle build). Stub of RayZath::Cuda::Engine for the CPU-only build of
// the reference: the constructor throws Cuda::Exception, so RayZath::Engine::Engine falls back to the
// CPU engine exactly as it does on a machine without CUDA (/root/reference/RayZath/rayzath.cpp:21-28).
#include "cuda_engine.cuh"
#include "cuda_exception.hpp"

namespace RayZath::Cuda
{
	class EngineCore {};

	Engine::Engine() { throw Exception("oracle build: CUDA engine not linked"); }
	Engine::~Engine() {}
	void Engine::renderWorld(RayZath::Engine::World&, const RayZath::Engine::RenderConfig&, const bool, const bool)
	{
		throw Exception("oracle build: CUDA engine not linked");
	}
	std::string Engine::timingsString() { return {}; }
}
