// Host BVH builders of the B200 render path (C ABI: rzb_build_mesh_bvh / rzb_build_instance_bvh).
//
// The render path consumes the reference's binary BVH as-is, because tree topology and leaf order
// decide which triangle wins an exact-t tie. When the caller is the
// reference's own World (the C++ drop-in, rayzath_b200/host/) the tree comes from the reference's
// host builder. When the caller has only triangles (Python host, bench, tests on the GPU box) this
// file builds the SAME tree: it follows the algorithm of
//   ComponentTreeNode::construct   /root/reference/RayZath/component_container.hpp:259-363 (triangles, leaf 8)
//   TreeNode::construct            /root/reference/RayZath/bvh_tree_node.hpp:117-215       (instances, leaf 4)
// and emits nodes/objects in the order of
//   Mesh::reconstruct              /root/reference/RayZath/cuda_instance.cu:161-220  (siblings adjacent, breadth pair first)
//   ObjectContainerWithBVH::constructNode  /root/reference/RayZath/cuda_bvh.cuh:86-111 (pre-order pairs)
// tests/test_flatten.py checks the output byte-for-byte against the reference builder (oracle/_ref).
//
// Floating-point notes: every operation is single precision in the reference's order; this file must
// be compiled with -ffp-contract=off (no FMA contraction) — see csrc/Makefile.

#include "../../include/rzb200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace
{
	struct V3
	{
		float x, y, z;
	};
	struct Box
	{
		V3 mn, mx;
	};

	inline V3 sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

	// BoundingBox(p1, p2): render_parts.cpp:157-168
	inline Box box2(const V3& a, const V3& b)
	{
		Box r;
		r.mn = {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)};
		r.mx = {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)};
		return r;
	}
	// BoundingBox::extendBy(point): render_parts.cpp:183-191
	inline void extend(Box& b, const V3& p)
	{
		if (b.mn.x > p.x) b.mn.x = p.x;
		if (b.mn.y > p.y) b.mn.y = p.y;
		if (b.mn.z > p.z) b.mn.z = p.z;
		if (b.mx.x < p.x) b.mx.x = p.x;
		if (b.mx.y < p.y) b.mx.y = p.y;
		if (b.mx.z < p.z) b.mx.z = p.z;
	}
	// BoundingBox::extendBy(box): render_parts.cpp:192-200
	inline void extend(Box& b, const Box& o)
	{
		if (b.mn.x > o.mn.x) b.mn.x = o.mn.x;
		if (b.mn.y > o.mn.y) b.mn.y = o.mn.y;
		if (b.mn.z > o.mn.z) b.mn.z = o.mn.z;
		if (b.mx.x < o.mx.x) b.mx.x = o.mx.x;
		if (b.mx.y < o.mx.y) b.mx.y = o.mx.y;
		if (b.mx.z < o.mx.z) b.mx.z = o.mx.z;
	}
	// BoundingBox::centroid: render_parts.cpp:201-204
	inline V3 centroid(const Box& b)
	{
		return {(b.mn.x + b.mx.x) * 0.5f, (b.mn.y + b.mx.y) * 0.5f, (b.mn.z + b.mx.z) * 0.5f};
	}

	// The libstdc++ and MSVC std::partition for bidirectional iterators are the same two-pointer sweep;
	// restated here so the object order does not depend on the standard library in use.
	template <class It, class Pred>
	It partition_ref(It first, It last, Pred pred)
	{
		for (;;)
		{
			for (;;)
			{
				if (first == last) return first;
				if (!pred(*first)) break;
				++first;
			}
			do
			{
				--last;
				if (first == last) return first;
			} while (!pred(*last));
			std::iter_swap(first, last);
			++first;
		}
	}

	enum SplitType : uint32_t { SplitZ = 0, SplitY = 1, SplitX = 2, SplitSize = 3 };

	struct BuildNode
	{
		Box bb;
		std::unique_ptr<BuildNode> first, second;
		SplitType type = SplitZ;
		uint32_t obj_begin = 0, obj_end = 0; // leaf: range in the (permuted) object id array
		bool leaf() const { return !first; }
		uint32_t treeSize() const { return leaf() ? 1u : first->treeSize() + second->treeSize() + 1u; }
	};

	struct Builder
	{
		const Box* boxes;   // per object
		std::vector<V3> centroids;
		uint32_t* ids;      // permuted in place
		uint32_t leaf_size, root_leaf_size, max_depth;

		std::unique_ptr<BuildNode> make(const Box& bb, uint32_t begin, uint32_t end, uint32_t depth)
		{
			auto node = std::make_unique<BuildNode>();
			node->bb = bb;
			construct(*node, begin, end, depth);
			fit(*node);
			return node;
		}
		void makeLeaf(BuildNode& n, uint32_t begin, uint32_t end)
		{
			n.obj_begin = begin;
			n.obj_end = end;
		}
		void fit(BuildNode& n)
		{
			Box bb{{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}}; // BoundingBox() default
			if (n.leaf())
			{
				if (n.obj_end > n.obj_begin)
				{
					bb = boxes[ids[n.obj_begin]];
					for (uint32_t i = n.obj_begin + 1; i < n.obj_end; ++i) extend(bb, boxes[ids[i]]);
				}
			}
			else
			{
				bb = n.first->bb;
				extend(bb, n.second->bb);
			}
			n.bb = bb;
		}
		void construct(BuildNode& n, uint32_t begin, uint32_t end, uint32_t depth)
		{
			const uint32_t count = end - begin;
			if (depth > max_depth || count <= leaf_size || (depth == 0 && count <= root_leaf_size))
			{
				makeLeaf(n, begin, end);
				return;
			}

			// objects not strictly smaller than the node on every axis are split off ("Size" partition)
			const V3 node_size = sub(n.bb.mx, n.bb.mn);
			uint32_t* const size_split = partition_ref(ids + begin, ids + end, [&](const uint32_t id) {
				const V3 s = sub(boxes[id].mx, boxes[id].mn);
				return s.x < node_size.x && s.y < node_size.y && s.z < node_size.z;
			});
			const uint32_t to_split_count = uint32_t(size_split - (ids + begin));
			const uint32_t too_large_count = count - to_split_count;
			const uint32_t split_end = begin + to_split_count;
			if (to_split_count != 0 && too_large_count != 0)
			{
				n.type = SplitSize;
				n.first = make(n.bb, begin, split_end, depth + 1);
				n.second = make(n.bb, split_end, end, depth + 1);
				return;
			}
			else if (to_split_count == 0)
			{
				makeLeaf(n, split_end, end);
				return;
			}

			// split point = running mean of centroids
			V3 sp{0.0f, 0.0f, 0.0f};
			for (uint32_t i = 0; i < to_split_count; ++i)
			{
				const V3& c = centroids[ids[begin + i]];
				const float d = float(int32_t(i) + 1);
				sp.x += (c.x - sp.x) / d;
				sp.y += (c.y - sp.y) / d;
				sp.z += (c.z - sp.z) / d;
			}
			// per-axis variance and "left of split" counts
			V3 var{0.0f, 0.0f, 0.0f};
			uint32_t cx = 0, cy = 0, cz = 0;
			for (uint32_t i = 0; i < to_split_count; ++i)
			{
				const V3& c = centroids[ids[begin + i]];
				const V3 d = sub(c, sp);
				var.x += d.x * d.x;
				var.y += d.y * d.y;
				var.z += d.z * d.z;
				cx += uint32_t(c.x < sp.x);
				cy += uint32_t(c.y < sp.y);
				cz += uint32_t(c.z < sp.z);
			}
			if (cx == 0 && cy == 0 && cz == 0)
			{
				makeLeaf(n, begin, split_end);
				return;
			}
			const float cnt = float(to_split_count);
			const V3 score{var.x / cnt, var.y / cnt, var.z / cnt};

			int axis;
			if (score.x >= score.y && score.x >= score.z && cx) axis = 0;
			else if (score.y >= score.x && score.y >= score.z && cy) axis = 1;
			else axis = 2;

			uint32_t* plane;
			V3 mx = n.bb.mx, mn = n.bb.mn;
			if (axis == 0)
			{
				plane = partition_ref(ids + begin, ids + split_end, [&](uint32_t id) { return centroids[id].x < sp.x; });
				mx.x = mn.x = sp.x;
				n.type = SplitX;
			}
			else if (axis == 1)
			{
				plane = partition_ref(ids + begin, ids + split_end, [&](uint32_t id) { return centroids[id].y < sp.y; });
				mx.y = mn.y = sp.y;
				n.type = SplitY;
			}
			else
			{
				plane = partition_ref(ids + begin, ids + split_end, [&](uint32_t id) { return centroids[id].z < sp.z; });
				mx.z = mn.z = sp.z;
				n.type = SplitZ;
			}
			const uint32_t mid = uint32_t(plane - ids);
			const Box parent = n.bb;
			n.first = make(box2(parent.mn, mx), begin, mid, depth + 1);
			n.second = make(box2(mn, parent.mx), mid, split_end, depth + 1);
		}
	};

	inline rzb_node makeNode(const Box& bb, uint32_t type, uint32_t begin, uint32_t count)
	{
		rzb_node n;
		n.bb_min[0] = bb.mn.x; n.bb_min[1] = bb.mn.y; n.bb_min[2] = bb.mn.z;
		n.bb_max[0] = bb.mx.x; n.bb_max[1] = bb.mx.y; n.bb_max[2] = bb.mx.z;
		n.begin = begin;
		n.type_count = ((type << 30) & 0xC0000000u) | (count & 0x3FFFFFFFu);
		return n;
	}

	// Mesh::reconstruct order (cuda_instance.cu:161-220)
	struct MeshFlattener
	{
		const uint32_t* ids;
		std::vector<rzb_node>& nodes;
		std::vector<uint32_t>& order;

		void addLeaf(const BuildNode& n)
		{
			nodes.push_back(makeNode(n.bb, 0, uint32_t(order.size()), n.obj_end - n.obj_begin));
			for (uint32_t i = n.obj_begin; i < n.obj_end; ++i) order.push_back(ids[i]);
		}
		void buildChildren(const BuildNode& n)
		{
			const BuildNode& c1 = *n.first;
			const uint32_t first_subtree = c1.treeSize() - 1;
			if (c1.leaf()) addLeaf(c1);
			else nodes.push_back(makeNode(c1.bb, c1.type, uint32_t(nodes.size()) + 2, 0));
			const BuildNode& c2 = *n.second;
			if (c2.leaf()) addLeaf(c2);
			else nodes.push_back(makeNode(c2.bb, c2.type, uint32_t(nodes.size()) + first_subtree + 1, 0));
			if (!c1.leaf()) buildChildren(c1);
			if (!c2.leaf()) buildChildren(c2);
		}
		void run(const BuildNode& root)
		{
			if (root.leaf()) addLeaf(root);
			else
			{
				nodes.push_back(makeNode(root.bb, root.type, uint32_t(nodes.size()) + 1, 0));
				buildChildren(root);
			}
		}
	};

	// ObjectContainerWithBVH::constructNode order (cuda_bvh.cuh:86-111)
	struct InstanceFlattener
	{
		const uint32_t* ids;
		std::vector<rzb_node>& nodes;
		std::vector<uint32_t>& order;

		void construct(size_t slot, const BuildNode& n)
		{
			if (n.leaf())
			{
				nodes[slot] = makeNode(n.bb, 0, uint32_t(order.size()), n.obj_end - n.obj_begin);
				for (uint32_t i = n.obj_begin; i < n.obj_end; ++i) order.push_back(ids[i]);
			}
			else
			{
				const size_t first = nodes.size();
				nodes[slot] = makeNode(n.bb, n.type, uint32_t(first), 0);
				nodes.emplace_back();
				nodes.emplace_back();
				construct(first, *n.first);
				construct(first + 1, *n.second);
			}
		}
		void run(const BuildNode& root)
		{
			nodes.emplace_back();
			construct(0, root);
		}
	};
}

extern "C" int rzb_build_mesh_bvh(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out)
{
	if (!vertices || !tris || !nodes_out || !node_count_out || !order_out) return RZB_ERR_INVALID;
	*node_count_out = 0;
	if (nt == 0) return RZB_OK;
	const V3* v = reinterpret_cast<const V3*>(vertices);

	std::vector<Box> boxes(nt);
	for (uint32_t i = 0; i < nt; ++i)
	{
		const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
		if (a >= nv || b >= nv || c >= nv) return RZB_ERR_INVALID;
		boxes[i] = box2(v[a], v[b]); // Triangle::boundingBox -> BoundingBox(v1, v2, v3), mesh_component.cpp:27-33
		extend(boxes[i], v[c]);
	}
	std::vector<uint32_t> ids(nt);
	for (uint32_t i = 0; i < nt; ++i) ids[i] = i;

	Builder b;
	b.boxes = boxes.data();
	b.centroids.resize(nt);
	for (uint32_t i = 0; i < nt; ++i) b.centroids[i] = centroid(boxes[i]);
	b.ids = ids.data();
	b.leaf_size = 8; b.root_leaf_size = 32; b.max_depth = 31; // component_container.hpp:145,265-267

	// ComponentTreeNode(mesh, components): bb = box of component 0 extended by every component (:183-195)
	Box root_bb = boxes[0];
	for (uint32_t i = 0; i < nt; ++i) extend(root_bb, boxes[i]);
	const auto root = b.make(root_bb, 0, nt, 0);

	std::vector<rzb_node> nodes;
	std::vector<uint32_t> order;
	nodes.reserve(root->treeSize());
	order.reserve(nt);
	MeshFlattener{ids.data(), nodes, order}.run(*root);
	if (nodes.size() > node_capacity || order.size() != nt) return RZB_ERR_NOMEM;
	std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(rzb_node));
	std::memcpy(order_out, order.data(), order.size() * sizeof(uint32_t));
	*node_count_out = uint32_t(nodes.size());
	return RZB_OK;
}

extern "C" int rzb_build_instance_bvh(const float* boxes_in, uint32_t n,
	rzb_node* nodes_out, uint32_t node_capacity, uint32_t* node_count_out, uint32_t* order_out)
{
	if (!nodes_out || !node_count_out || (n && (!boxes_in || !order_out))) return RZB_ERR_INVALID;
	*node_count_out = 0;
	std::vector<Box> boxes(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		boxes[i].mn = {boxes_in[6 * i], boxes_in[6 * i + 1], boxes_in[6 * i + 2]};
		boxes[i].mx = {boxes_in[6 * i + 3], boxes_in[6 * i + 4], boxes_in[6 * i + 5]};
	}
	std::vector<uint32_t> ids(n);
	for (uint32_t i = 0; i < n; ++i) ids[i] = i;

	Builder b;
	b.boxes = boxes.data();
	b.centroids.resize(n);
	for (uint32_t i = 0; i < n; ++i) b.centroids[i] = centroid(boxes[i]);
	b.ids = ids.data();
	b.leaf_size = 4; b.root_leaf_size = 8; b.max_depth = 31; // bvh_tree_node.hpp:14,120-121

	// ObjectContainerWithBVH::update (bvh.hpp:29-53): bb starts at object 0's box, extended by all
	Box root_bb{{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
	if (n) root_bb = boxes[0];
	for (uint32_t i = 0; i < n; ++i) extend(root_bb, boxes[i]);
	const auto root = b.make(root_bb, 0, n, 0);

	std::vector<rzb_node> nodes;
	std::vector<uint32_t> order;
	InstanceFlattener{ids.data(), nodes, order}.run(*root);
	if (nodes.size() > node_capacity) return RZB_ERR_NOMEM;
	std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(rzb_node));
	if (n) std::memcpy(order_out, order.data(), order.size() * sizeof(uint32_t));
	*node_count_out = uint32_t(nodes.size());
	return RZB_OK;
}
