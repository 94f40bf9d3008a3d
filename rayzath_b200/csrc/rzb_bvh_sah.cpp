// Optional triangle-BVH builder of the B200 render path: binned surface-area-heuristic splits,
// emitted in the SAME node / triangle-order format as the reference's tree (rzb_node, include/rzb200.h), so the
// traversal kernels run on it unchanged. It is NOT the reference's tree (component_container.hpp:259-363 splits at the
// mean centroid of the axis of largest variance): closest-hit records are equal to the reference's except on
// exact-distance ties, where the winner depends on visiting order (as it already does between the reference's own CPU
// and CUDA engines); the reference tree stays the parity mode.
#include "../../include/rzb200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace
{
	struct V3 { float x, y, z; };
	struct Box
	{
		V3 mn{FLT_MAX, FLT_MAX, FLT_MAX}, mx{-FLT_MAX, -FLT_MAX, -FLT_MAX};
		void add(const V3& p)
		{
			mn.x = std::min(mn.x, p.x); mn.y = std::min(mn.y, p.y); mn.z = std::min(mn.z, p.z);
			mx.x = std::max(mx.x, p.x); mx.y = std::max(mx.y, p.y); mx.z = std::max(mx.z, p.z);
		}
		void add(const Box& b) { add(b.mn); add(b.mx); }
		float area() const
		{
			const float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
			return (dx < 0.0f) ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx);
		}
	};
	inline float axis(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

	constexpr int kBins = 32;
	// cost of one sibling-pair step relative to one triangle test (both ~60-80 instructions in k_trace_paths)
	constexpr float kPairCost = 1.5f; // measured best of 0.5 .. 4 on both bench scenes
	inline float pairCost()
	{
		// RZB200_SAH_PAIR_COST: tuning aid
		static const float c = [] { const char* e = std::getenv("RZB200_SAH_PAIR_COST"); return e ? float(std::atof(e)) : kPairCost; }();
		return c;
	}

	struct Task { uint32_t node, begin, end, depth; };
	constexpr uint32_t kMaxDepth = 31; // as the reference's trees (component_container.hpp:145): bounds the traversal stack
}

extern "C" int rzb_build_mesh_bvh_sah(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	uint32_t max_leaf, rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out)
{
	if (!vertices || !tris || !nodes_out || !node_count_out || !order_out) return RZB_ERR_INVALID;
	*node_count_out = 0;
	if (nt == 0) return RZB_OK;
	max_leaf = std::max(1u, std::min(max_leaf, 0x3FFFFFFFu));
	const V3* v = reinterpret_cast<const V3*>(vertices);

	std::vector<Box> boxes(nt);
	std::vector<V3> cent(nt);
	for (uint32_t i = 0; i < nt; ++i)
	{
		const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
		if (a >= nv || b >= nv || c >= nv) return RZB_ERR_INVALID;
		boxes[i].add(v[a]); boxes[i].add(v[b]); boxes[i].add(v[c]);
		cent[i] = {0.5f * (boxes[i].mn.x + boxes[i].mx.x), 0.5f * (boxes[i].mn.y + boxes[i].mx.y), 0.5f * (boxes[i].mn.z + boxes[i].mx.z)};
	}
	std::vector<uint32_t> ids(nt);
	for (uint32_t i = 0; i < nt; ++i) ids[i] = i;

	// nodes are allocated as sibling pairs from an atomic cursor (pairs start at odd indices, as the kernels' 64-byte
	// pair fetch wants); big ranges go through a shared queue served by all host threads, small ones stay on the
	// worker's own stack
	std::vector<rzb_node> nodes(2 * size_t(nt) + 1);
	std::atomic<uint32_t> cursor{1u};
	std::mutex mtx;
	std::condition_variable cv;
	std::deque<Task> shared{{0u, 0u, nt, 0u}};
	uint32_t busy = 0;
	const uint32_t kShareAbove = 8192;

	auto process = [&](const Task& t, std::vector<Task>& local) {
		const uint32_t n = t.end - t.begin;
		Box bb, cb;
		for (uint32_t i = t.begin; i < t.end; ++i) { bb.add(boxes[ids[i]]); cb.add(cent[ids[i]]); }
		rzb_node node{};
		node.bb_min[0] = bb.mn.x; node.bb_min[1] = bb.mn.y; node.bb_min[2] = bb.mn.z;
		node.bb_max[0] = bb.mx.x; node.bb_max[1] = bb.mx.y; node.bb_max[2] = bb.mx.z;

		// best binned split over the three axes
		float best_cost = FLT_MAX;
		int best_axis = -1, best_bin = 0;
		if (n > 1)
		{
			for (int a = 0; a < 3; ++a)
			{
				const float lo = axis(cb.mn, a), hi = axis(cb.mx, a);
				if (!(hi > lo)) continue;
				const float scale = float(kBins) / (hi - lo);
				Box bin_box[kBins];
				uint32_t bin_n[kBins] = {};
				for (uint32_t i = t.begin; i < t.end; ++i)
				{
					const uint32_t id = ids[i];
					const int b = std::min(kBins - 1, std::max(0, int((axis(cent[id], a) - lo) * scale)));
					bin_box[b].add(boxes[id]);
					bin_n[b]++;
				}
				float right_area[kBins];
				uint32_t right_n[kBins];
				Box acc;
				uint32_t cnt = 0;
				for (int b = kBins - 1; b > 0; --b)
				{
					if (bin_n[b]) acc.add(bin_box[b]);
					cnt += bin_n[b];
					right_area[b] = acc.area(); right_n[b] = cnt;
				}
				acc = Box{}; cnt = 0;
				for (int b = 0; b < kBins - 1; ++b)
				{
					if (bin_n[b]) acc.add(bin_box[b]);
					cnt += bin_n[b];
					if (cnt == 0 || right_n[b + 1] == 0) continue;
					const float cost = acc.area() * float(cnt) + right_area[b + 1] * float(right_n[b + 1]);
					if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
				}
			}
		}
		const float parent_area = bb.area();
		const float split_cost = (best_axis >= 0 && parent_area > 0.0f) ? pairCost() + best_cost / parent_area : FLT_MAX;
		// past depth 24 splits are at the median, so the depth limit is only reached by > 2^7 * max_leaf coincident triangles
		if (t.depth >= 24u) best_axis = -1;
		const bool leaf = n == 1 || t.depth >= kMaxDepth || (n <= max_leaf && float(n) <= split_cost);
		if (leaf)
		{
			node.begin = t.begin;
			node.type_count = n; // split type bits unused for leaves
			nodes[t.node] = node;
			return;
		}
		uint32_t mid;
		int split_axis = best_axis;
		if (best_axis >= 0)
		{
			const float lo = axis(cb.mn, best_axis), hi = axis(cb.mx, best_axis);
			const float scale = float(kBins) / (hi - lo);
			uint32_t* first = ids.data() + t.begin;
			uint32_t* last = ids.data() + t.end;
			uint32_t* m = std::partition(first, last, [&](uint32_t id) {
				const int b = std::min(kBins - 1, std::max(0, int((axis(cent[id], best_axis) - lo) * scale)));
				return b <= best_bin;
			});
			mid = uint32_t(m - ids.data());
		}
		else
		{
			// all centroids coincide: split the index range in half
			mid = t.begin + n / 2;
			split_axis = 0;
		}
		if (mid == t.begin || mid == t.end) mid = t.begin + n / 2;
		const uint32_t child = cursor.fetch_add(2u);
		
		node.begin = child;
		// split type in bits 30..31: X = 2, Y = 1, Z = 0 (bvh_tree_node.hpp:21-27); first child = lower side
		node.type_count = uint32_t(2 - split_axis) << 30;
		nodes[t.node] = node;
		const Task a{child, t.begin, mid, t.depth + 1u}, b{child + 1u, mid, t.end, t.depth + 1u};
		for (const Task& c : {b, a})
		{
			if (c.end - c.begin > kShareAbove)
			{
				std::lock_guard<std::mutex> lg(mtx);
				shared.push_back(c);
				cv.notify_one();
			}
			else local.push_back(c);
		}
	};
	auto worker = [&]() {
		std::vector<Task> local;
		for (;;)
		{
			Task t;
			{
				std::unique_lock<std::mutex> lk(mtx);
				cv.wait(lk, [&] { return !shared.empty() || busy == 0; });
				if (shared.empty()) { cv.notify_all(); return; }
				t = shared.front();
				shared.pop_front();
				++busy;
			}
			local.push_back(t);
			while (!local.empty())
			{
				const Task c = local.back();
				local.pop_back();
				process(c, local);
			}
			{
				std::lock_guard<std::mutex> lg(mtx);
				--busy;
			}
			cv.notify_all();
		}
	};
	const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	const unsigned n_threads = nt < 4 * kShareAbove ? 1u : hw;
	std::vector<std::thread> pool;
	for (unsigned i = 1; i < n_threads; ++i) pool.emplace_back(worker);
	worker();
	for (auto& th : pool) th.join();
	nodes.resize(cursor.load());
	{
		// canonical layout, independent of thread timing: root, then per inner node its sibling pair followed by the
		// first child's subtree and the second child's (the order Mesh::reconstruct uses, cuda_instance.cu:161-220)
		std::vector<rzb_node> laid(nodes.size());
		laid[0] = nodes[0];
		uint32_t next = 1u;
		std::vector<uint32_t> todo{0u}; // indices into `laid` whose children still live at old indices
		while (!todo.empty())
		{
			const uint32_t i = todo.back();
			todo.pop_back();
			if ((laid[i].type_count & 0x3FFFFFFFu) != 0u) continue;
			const uint32_t old_child = laid[i].begin;
			laid[next] = nodes[old_child];
			laid[next + 1u] = nodes[old_child + 1u];
			laid[i].begin = next;
			todo.push_back(next + 1u);
			todo.push_back(next);
			next += 2u;
		}
		nodes.swap(laid);
	}
	if (nodes.size() > node_capacity) return RZB_ERR_NOMEM;
	std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(rzb_node));
	std::memcpy(order_out, ids.data(), size_t(nt) * sizeof(uint32_t));
	*node_count_out = uint32_t(nodes.size());
	return RZB_OK;
}
