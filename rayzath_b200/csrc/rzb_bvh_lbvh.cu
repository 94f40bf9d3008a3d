// GPU triangle-BVH builder of the B200 render path ("GPU BVH build"): a linear BVH
// (Morton codes of the triangle centroids -> radix sort -> Karras' parallel radix tree -> bottom-up boxes), leaves
// collapsed to <= max_leaf triangles, emitted in the SAME node / triangle-order format as the reference's tree
// (rzb_node, include/rzb200.h: children as adjacent pairs at odd indices), so every kernel runs on it unchanged
// (RZB_SCENE_OWN_TREES). Built for (re)build SPEED -- 1M triangles in 2.5 ms of device time against
// ~0.5 s for either host builder; the SAH tree of rzb_build_mesh_bvh_sah traces faster.
// The sort is CUB's device radix sort (library plumbing); tree construction, boxes, collapse and emission are the
// kernels below. Not the reference's tree: hit records equal the reference's except exact-distance ties.
#include "../../include/rzb200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <utility>
#include <vector>

namespace
{
	struct Box3 { float mnx, mny, mnz, mxx, mxy, mxz; };

	__device__ __forceinline__ uint32_t expand10(uint32_t v)
	{
		v = (v * 0x00010001u) & 0xFF0000FFu;
		v = (v * 0x00000101u) & 0x0F00F00Fu;
		v = (v * 0x00000011u) & 0xC30C30C3u;
		v = (v * 0x00000005u) & 0x49249249u;
		return v;
	}

	// per triangle: box and 64-bit key = 30-bit Morton code of the box centre (within the mesh box) << 32 | index
	__global__ void k_lbvh_keys(const float* __restrict__ v, const uint32_t* __restrict__ tris, uint32_t nt,
		Box3 scene, Box3* __restrict__ tri_box, unsigned long long* __restrict__ keys)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= nt) return;
		const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
		const float ax = v[3 * a], ay = v[3 * a + 1], az = v[3 * a + 2];
		const float bx = v[3 * b], by = v[3 * b + 1], bz = v[3 * b + 2];
		const float cx = v[3 * c], cy = v[3 * c + 1], cz = v[3 * c + 2];
		Box3 t;
		t.mnx = fminf(ax, fminf(bx, cx)); t.mny = fminf(ay, fminf(by, cy)); t.mnz = fminf(az, fminf(bz, cz));
		t.mxx = fmaxf(ax, fmaxf(bx, cx)); t.mxy = fmaxf(ay, fmaxf(by, cy)); t.mxz = fmaxf(az, fmaxf(bz, cz));
		tri_box[i] = t;
		// one scale for all axes (the largest extent): a cell is a cube, so a flat mesh does not spend key bits on
		// splitting its thin axis
		const float s = fmaxf(fmaxf(scene.mxx - scene.mnx, scene.mxy - scene.mny), scene.mxz - scene.mnz);
		const float fx = s > 0.0f ? (0.5f * (t.mnx + t.mxx) - scene.mnx) / s : 0.0f;
		const float fy = s > 0.0f ? (0.5f * (t.mny + t.mxy) - scene.mny) / s : 0.0f;
		const float fz = s > 0.0f ? (0.5f * (t.mnz + t.mxz) - scene.mnz) / s : 0.0f;
		const uint32_t qx = min(1023u, uint32_t(fmaxf(fx, 0.0f) * 1024.0f));
		const uint32_t qy = min(1023u, uint32_t(fmaxf(fy, 0.0f) * 1024.0f));
		const uint32_t qz = min(1023u, uint32_t(fmaxf(fz, 0.0f) * 1024.0f));
		const uint32_t morton = (expand10(qx) << 2) | (expand10(qy) << 1) | expand10(qz);
		keys[i] = ((unsigned long long)morton << 32) | i;
	}

	// Karras 2012: internal node i of the radix tree over the sorted, unique keys. Children >= nt - 1 + ... are
	// encoded as: internal j -> j, leaf k -> k | kLeaf.
	constexpr uint32_t kLeaf = 0x80000000u;
	constexpr uint32_t kMaxTreeDepth = 38; // top level (<= 31) + 1 + this + slack fits the 72-entry traversal stack
	__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j)
	{
		if (j < 0 || j >= n) return -1;
		return __clzll(keys[i] ^ keys[j]);
	}
	__global__ void k_lbvh_tree(const unsigned long long* __restrict__ keys, int n, uint32_t* __restrict__ left,
		uint32_t* __restrict__ right, uint32_t* __restrict__ first, uint32_t* __restrict__ last,
		uint32_t* __restrict__ parent_internal, uint32_t* __restrict__ parent_leaf)
	{
		const int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n - 1) return;
		const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
		const int dmin = delta(keys, n, i, i - d);
		int lmax = 2;
		while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
		int l = 0;
		for (int t = lmax / 2; t >= 1; t /= 2)
			if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
		const int j = i + l * d;
		const int dnode = delta(keys, n, i, j);
		int s = 0;
		for (int t = (l + 1) / 2;; t = (t + 1) / 2)
		{
			if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
			if (t == 1) break;
		}
		const int gamma = i + s * d + min(d, 0);
		const int lo = min(i, j), hi = max(i, j);
		const uint32_t lc = (lo == gamma) ? (uint32_t(gamma) | kLeaf) : uint32_t(gamma);
		const uint32_t rc = (hi == gamma + 1) ? (uint32_t(gamma + 1) | kLeaf) : uint32_t(gamma + 1);
		left[i] = lc; right[i] = rc;
		first[i] = uint32_t(lo); last[i] = uint32_t(hi);
		if (lc & kLeaf) parent_leaf[gamma] = uint32_t(i); else parent_internal[gamma] = uint32_t(i);
		if (rc & kLeaf) parent_leaf[gamma + 1] = uint32_t(i); else parent_internal[gamma + 1] = uint32_t(i);
		if (i == 0) parent_internal[0] = 0xFFFFFFFFu;
	}

	__device__ __forceinline__ Box3 unite(const Box3& a, const Box3& b)
	{
		Box3 r;
		r.mnx = fminf(a.mnx, b.mnx); r.mny = fminf(a.mny, b.mny); r.mnz = fminf(a.mnz, b.mnz);
		r.mxx = fmaxf(a.mxx, b.mxx); r.mxy = fmaxf(a.mxy, b.mxy); r.mxz = fmaxf(a.mxz, b.mxz);
		return r;
	}

	// boxes written by other thread blocks during k_lbvh_boxes are read past L1 (a line may have been cached before its
	// neighbour in the same line was written)
	__device__ __forceinline__ Box3 load_box_cg(const Box3* p)
	{
		const float* f = reinterpret_cast<const float*>(p);
		Box3 b;
		b.mnx = __ldcg(f); b.mny = __ldcg(f + 1); b.mnz = __ldcg(f + 2);
		b.mxx = __ldcg(f + 3); b.mxy = __ldcg(f + 4); b.mxz = __ldcg(f + 5);
		return b;
	}

	// bottom-up: the second thread to arrive at an internal node unites its children's boxes and moves on
	__global__ void k_lbvh_boxes(const unsigned long long* __restrict__ keys, int n, const Box3* __restrict__ tri_box,
		const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const uint32_t* __restrict__ parent_internal,
		const uint32_t* __restrict__ parent_leaf, Box3* node_box, uint32_t* __restrict__ arrived)
	{
		const int k = blockIdx.x * blockDim.x + threadIdx.x;
		if (k >= n) return;
		uint32_t p = parent_leaf[k];
		while (p != 0xFFFFFFFFu)
		{
			__threadfence();
			if (atomicAdd(arrived + p, 1u) == 0u) return; // the sibling subtree is not finished yet
			const uint32_t lc = left[p], rc = right[p];
			const Box3 a = (lc & kLeaf) ? tri_box[uint32_t(keys[lc & ~kLeaf])] : load_box_cg(node_box + lc);
			const Box3 b = (rc & kLeaf) ? tri_box[uint32_t(keys[rc & ~kLeaf])] : load_box_cg(node_box + rc);
			node_box[p] = unite(a, b);
			p = parent_internal[p];
		}
	}

	// an internal node survives as an inner node iff its range holds more than max_leaf triangles
	__global__ void k_lbvh_flags(int n, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last,
		uint32_t max_leaf, uint32_t* __restrict__ flag)
	{
		const int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n - 1) return;
		flag[i] = (last[i] - first[i] + 1u > max_leaf) ? 1u : 0u;
	}

	__device__ __forceinline__ rzb_node make_node(const Box3& b, uint32_t begin, uint32_t type_count)
	{
		rzb_node r;
		r.bb_min[0] = b.mnx; r.bb_min[1] = b.mny; r.bb_min[2] = b.mnz;
		r.bb_max[0] = b.mxx; r.bb_max[1] = b.mxy; r.bb_max[2] = b.mxz;
		r.begin = begin; r.type_count = type_count;
		return r;
	}

	// surviving inner node i (compact index p) writes its two children at nodes[1 + 2p], nodes[2 + 2p]
	__global__ void k_lbvh_emit(const unsigned long long* __restrict__ keys, int n, const Box3* __restrict__ tri_box,
		const Box3* __restrict__ node_box, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
		const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, const uint32_t* __restrict__ flag,
		const uint32_t* __restrict__ slot, rzb_node* __restrict__ nodes)
	{
		const int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n - 1 || !flag[i]) return;
		const uint32_t p = slot[i];
		if (i == 0) nodes[0] = make_node(node_box[0], 1u, 0u);
		const uint32_t child[2] = {left[i], right[i]};
#pragma unroll
		for (int k = 0; k < 2; ++k)
		{
			const uint32_t c = child[k];
			rzb_node out;
			if (c & kLeaf)
			{
				const uint32_t pos = c & ~kLeaf;
				out = make_node(tri_box[uint32_t(keys[pos])], pos, 1u);
			}
			else if (!flag[c]) out = make_node(node_box[c], first[c], last[c] - first[c] + 1u);
			else out = make_node(node_box[c], 1u + 2u * slot[c], 0u);
			nodes[1u + 2u * p + uint32_t(k)] = out;
		}
	}

	__global__ void k_lbvh_order(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ order)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i < n) order[i] = uint32_t(keys[i]);
	}

	struct DeviceArena
	{
		std::vector<void*> ptrs;
		~DeviceArena() { for (void* p : ptrs) cudaFree(p); }
		template <typename T> T* get(size_t count)
		{
			void* p = nullptr;
			if (cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return nullptr;
			ptrs.push_back(p);
			return static_cast<T*>(p);
		}
	};
}

extern "C" int rzb_build_mesh_bvh_lbvh(int device, const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	uint32_t max_leaf, rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out,
	float* device_ms_out)
{
	if (!vertices || !tris || !nodes_out || !node_count_out || !order_out) return RZB_ERR_INVALID;
	*node_count_out = 0;
	if (device_ms_out) *device_ms_out = 0.0f;
	if (nt == 0) return RZB_OK;
	max_leaf = std::max(1u, std::min(max_leaf, 0x3FFFFFFFu));
	for (size_t i = 0; i < size_t(nt) * 3; ++i)
		if (tris[i] >= nv) return RZB_ERR_INVALID;
	Box3 scene{FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (uint32_t i = 0; i < nv; ++i)
	{
		scene.mnx = std::min(scene.mnx, vertices[3 * i]); scene.mxx = std::max(scene.mxx, vertices[3 * i]);
		scene.mny = std::min(scene.mny, vertices[3 * i + 1]); scene.mxy = std::max(scene.mxy, vertices[3 * i + 1]);
		scene.mnz = std::min(scene.mnz, vertices[3 * i + 2]); scene.mxz = std::max(scene.mxz, vertices[3 * i + 2]);
	}
	int previous = 0;
	if (cudaGetDevice(&previous) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return RZB_ERR_CUDA;
	struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{previous};

	const int n = int(nt);
	if (n <= int(max_leaf))
	{
		// one leaf: no tree to build
		Box3 b{FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
		for (uint32_t i = 0; i < nt; ++i)
			for (int k = 0; k < 3; ++k)
			{
				const float* p = vertices + 3 * size_t(tris[3 * i + k]);
				b.mnx = std::min(b.mnx, p[0]); b.mny = std::min(b.mny, p[1]); b.mnz = std::min(b.mnz, p[2]);
				b.mxx = std::max(b.mxx, p[0]); b.mxy = std::max(b.mxy, p[1]); b.mxz = std::max(b.mxz, p[2]);
			}
		if (node_capacity < 1) return RZB_ERR_NOMEM;
		rzb_node r{};
		r.bb_min[0] = b.mnx; r.bb_min[1] = b.mny; r.bb_min[2] = b.mnz;
		r.bb_max[0] = b.mxx; r.bb_max[1] = b.mxy; r.bb_max[2] = b.mxz;
		r.begin = 0; r.type_count = nt;
		nodes_out[0] = r;
		for (uint32_t i = 0; i < nt; ++i) order_out[i] = i;
		*node_count_out = 1;
		return RZB_OK;
	}

	DeviceArena mem;
	float* d_v = mem.get<float>(size_t(nv) * 3);
	uint32_t* d_t = mem.get<uint32_t>(size_t(nt) * 3);
	Box3* d_tri_box = mem.get<Box3>(nt);
	unsigned long long* d_keys = mem.get<unsigned long long>(nt);
	unsigned long long* d_keys_sorted = mem.get<unsigned long long>(nt);
	uint32_t* d_left = mem.get<uint32_t>(nt);
	uint32_t* d_right = mem.get<uint32_t>(nt);
	uint32_t* d_first = mem.get<uint32_t>(nt);
	uint32_t* d_last = mem.get<uint32_t>(nt);
	uint32_t* d_parent_internal = mem.get<uint32_t>(nt);
	uint32_t* d_parent_leaf = mem.get<uint32_t>(nt);
	uint32_t* d_arrived = mem.get<uint32_t>(nt);
	uint32_t* d_flag = mem.get<uint32_t>(nt);
	uint32_t* d_slot = mem.get<uint32_t>(nt);
	Box3* d_node_box = mem.get<Box3>(nt);
	uint32_t* d_order = mem.get<uint32_t>(nt);
	if (!d_v || !d_t || !d_tri_box || !d_keys || !d_keys_sorted || !d_left || !d_right || !d_first || !d_last ||
		!d_parent_internal || !d_parent_leaf || !d_arrived || !d_flag || !d_slot || !d_node_box || !d_order)
		return RZB_ERR_NOMEM;
	size_t sort_bytes = 0, scan_bytes = 0;
	cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, d_keys, d_keys_sorted, n, 0, 62);
	cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flag, d_slot, n - 1);
	void* d_temp = mem.get<uint8_t>(std::max(sort_bytes, scan_bytes));
	if (!d_temp) return RZB_ERR_NOMEM;

	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return RZB_ERR_CUDA;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaMemcpyAsync(d_v, vertices, size_t(nv) * 12, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(d_t, tris, size_t(nt) * 12, cudaMemcpyHostToDevice, stream);
	cudaEventRecord(e0, stream);
	const int B = 256, G = (n + B - 1) / B;
	k_lbvh_keys<<<G, B, 0, stream>>>(d_v, d_t, nt, scene, d_tri_box, d_keys);
	size_t temp_bytes = std::max(sort_bytes, scan_bytes);
	cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, d_keys, d_keys_sorted, n, 0, 62, stream);
	cudaMemsetAsync(d_arrived, 0, size_t(nt) * 4, stream);
	k_lbvh_tree<<<G, B, 0, stream>>>(d_keys_sorted, n, d_left, d_right, d_first, d_last, d_parent_internal, d_parent_leaf);
	k_lbvh_boxes<<<G, B, 0, stream>>>(d_keys_sorted, n, d_tri_box, d_left, d_right, d_parent_internal, d_parent_leaf, d_node_box, d_arrived);
	k_lbvh_flags<<<G, B, 0, stream>>>(n, d_first, d_last, max_leaf, d_flag);
	temp_bytes = std::max(sort_bytes, scan_bytes);
	cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_flag, d_slot, n - 1, stream);
	uint32_t last_flag = 0, last_slot = 0;
	cudaMemcpyAsync(&last_flag, d_flag + (n - 2), 4, cudaMemcpyDeviceToHost, stream);
	cudaMemcpyAsync(&last_slot, d_slot + (n - 2), 4, cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	const uint32_t inner = last_slot + last_flag; // surviving inner nodes (>= 1: the root holds n > max_leaf triangles)
	const uint32_t node_count = 1u + 2u * inner;
	int rc = RZB_OK;
	if (node_count > node_capacity) rc = RZB_ERR_NOMEM;
	rzb_node* d_nodes = rc == RZB_OK ? mem.get<rzb_node>(node_count) : nullptr;
	if (rc == RZB_OK && !d_nodes) rc = RZB_ERR_NOMEM;
	if (rc == RZB_OK)
	{
		k_lbvh_emit<<<G, B, 0, stream>>>(d_keys_sorted, n, d_tri_box, d_node_box, d_left, d_right, d_first, d_last, d_flag, d_slot, d_nodes);
		k_lbvh_order<<<G, B, 0, stream>>>(d_keys_sorted, nt, d_order);
		cudaEventRecord(e1, stream);
		cudaMemcpyAsync(nodes_out, d_nodes, size_t(node_count) * sizeof(rzb_node), cudaMemcpyDeviceToHost, stream);
		cudaMemcpyAsync(order_out, d_order, size_t(nt) * 4, cudaMemcpyDeviceToHost, stream);
		if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = RZB_ERR_CUDA;
		else
		{
			*node_count_out = node_count;
			if (device_ms_out) cudaEventElapsedTime(device_ms_out, e0, e1);
		}
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaStreamDestroy(stream);
	if (rc == RZB_OK)
	{
		// The traversal stack is sized for trees of depth <= 31 per level like the reference's; a radix tree over
		// clustered centroids can be deeper (up to one level per key bit). Such inputs go to the SAH builder.
		uint32_t max_depth = 0;
		std::vector<std::pair<uint32_t, uint32_t>> todo{{0u, 0u}};
		while (!todo.empty())
		{
			const auto [i, depth] = todo.back();
			todo.pop_back();
			max_depth = std::max(max_depth, depth);
			if ((nodes_out[i].type_count & 0x3FFFFFFFu) == 0u)
			{
				todo.push_back({nodes_out[i].begin, depth + 1u});
				todo.push_back({nodes_out[i].begin + 1u, depth + 1u});
			}
		}
		if (max_depth > kMaxTreeDepth)
			return rzb_build_mesh_bvh_sah(vertices, nv, tris, nt, max_leaf, nodes_out, node_capacity, node_count_out, order_out);
	}
	return rc;
}
