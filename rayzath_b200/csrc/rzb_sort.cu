// Prefix sum over the ray-order bins (rzb_render): a thin wrapper around CUB's DeviceScan so that rzb_api.cu does not
// have to compile CUB.
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rzb
{
	size_t scanTempBytes(uint32_t n)
	{
		size_t bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(n));
		return bytes;
	}
	cudaError_t exclusiveScan(void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, uint32_t n, cudaStream_t stream)
	{
		return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, int(n), stream);
	}
}
