// TEST INFRASTRUCTURE (oracle): closest-hit records straight from the REFERENCE'S OWN CUDA traversal.
// One thread per ray runs what World::closestObjectIntersection / World::rayCast run first
// (/root/reference/RayZath/cuda_world.cuh:80-90, 105-115: `instances.closestIntersection(ray, traversal)` on the
// device World the reference engine mirrored itself) and writes TraversalResult (cuda_render_parts.cuh:946-952) as plain
// numbers. Compiled with the reference's cuda_*.cu files (oracle/Makefile, target ref_cuda) into rz_ref_tool_cuda;
// nothing of it is linked into the product.
#define private public
#define protected public
#include "cuda_engine.cuh"
#include "cuda_engine_core.cuh"
#include "cuda_world.cuh"
#undef private
#undef protected

#include "ref_trace_cuda.h"

namespace
{
	using namespace RayZath::Cuda;

	__global__ void refTraceKernel(World* world, const float* __restrict__ o, const float* __restrict__ d,
		const float* __restrict__ nf, uint32_t n, RefCudaHit* __restrict__ out)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		RangedRay ray(vec3f(o[3 * i], o[3 * i + 1], o[3 * i + 2]), vec3f(d[3 * i], d[3 * i + 1], d[3 * i + 2]),
			vec2f(nf[2 * i], nf[2 * i + 1]));
		TraversalResult traversal;
		world->instances.closestIntersection(ray, traversal);
		RefCudaHit h;
		h.instance = 0xFFFFFFFFu; h.triangle_bvh_order = 0xFFFFFFFFu; h.b1 = 0.0f; h.b2 = 0.0f; h.external = 0u;
		h.t = ray.near_far.y;
		if (traversal.closest_instance)
		{
			h.instance = traversal.closest_instance->m_instance_idx;
			if (traversal.closest_triangle && traversal.closest_instance->mesh)
				h.triangle_bvh_order = uint32_t(traversal.closest_triangle - traversal.closest_instance->mesh->mp_triangles);
			h.b1 = traversal.barycenter.x; h.b2 = traversal.barycenter.y;
			h.external = traversal.external ? 1u : 0u;
		}
		out[i] = h;
	}
}

int refCudaTrace(void* cuda_engine, const float* origins, const float* directions, const float* near_far, uint32_t n,
	RefCudaHit* hits_out)
{
	auto* core = static_cast<RayZath::Cuda::Engine*>(cuda_engine)->m_engine_core.get();
	World* d_world = core->cudaWorld();
	float *d_o = nullptr, *d_d = nullptr, *d_nf = nullptr;
	RefCudaHit* d_hits = nullptr;
	if (cudaMalloc(&d_o, size_t(n) * 12) || cudaMalloc(&d_d, size_t(n) * 12) || cudaMalloc(&d_nf, size_t(n) * 8) ||
		cudaMalloc(&d_hits, size_t(n) * sizeof(RefCudaHit))) return 1;
	cudaMemcpy(d_o, origins, size_t(n) * 12, cudaMemcpyHostToDevice);
	cudaMemcpy(d_d, directions, size_t(n) * 12, cudaMemcpyHostToDevice);
	cudaMemcpy(d_nf, near_far, size_t(n) * 8, cudaMemcpyHostToDevice);
	refTraceKernel<<<(n + 127) / 128, 128>>>(d_world, d_o, d_d, d_nf, n, d_hits);
	const cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { std::fprintf(stderr, "refCudaTrace: %s\n", cudaGetErrorString(e)); return 2; }
	cudaMemcpy(hits_out, d_hits, size_t(n) * sizeof(RefCudaHit), cudaMemcpyDeviceToHost);
	cudaFree(d_o); cudaFree(d_d); cudaFree(d_nf); cudaFree(d_hits);
	return 0;
}
