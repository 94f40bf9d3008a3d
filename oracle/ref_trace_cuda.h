// TEST INFRASTRUCTURE (oracle): see ref_trace_cuda.cu
#pragma once
#include <stdint.h>

struct RefCudaHit
{
	uint32_t instance;           // Instance::m_instance_idx (host container index), 0xFFFFFFFF = miss
	uint32_t triangle_bvh_order; // index into the device mesh's triangle array (= the order Mesh::reconstruct emitted)
	float t;                     // ray.near_far.y after traversal
	float b1, b2;
	uint32_t external;
};

// cuda_engine: RayZath::Cuda::Engine* (Engine::m_cuda_engine) after it has mirrored the world (one renderWorld call)
int refCudaTrace(void* cuda_engine, const float* origins, const float* directions, const float* near_far, uint32_t n,
	RefCudaHit* hits_out);
