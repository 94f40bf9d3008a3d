// Key-value radix sort of the ray-order pass (rzb_render with ray sorting on): a thin wrapper around CUB's
// DeviceRadixSort so that rzb_api.cu does not have to compile CUB. Keys are the 24/25-bit ray keys k_shade writes
// (ray class | Morton cell of the origin | direction octant), values the slot indices.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rzb
{
	size_t sortTempBytes(uint32_t n, int end_bit)
	{
		size_t bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
			static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(n), 0, end_bit);
		return bytes;
	}
	cudaError_t sortPairs(void* temp, size_t temp_bytes, const uint32_t* keys_in, uint32_t* keys_out, const uint32_t* vals_in,
		uint32_t* vals_out, uint32_t n, int end_bit, cudaStream_t stream)
	{
		return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, int(n), 0, end_bit, stream);
	}
}
