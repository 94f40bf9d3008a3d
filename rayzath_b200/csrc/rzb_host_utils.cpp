// Host-side float utilities of the C ABI that must round exactly like the reference's host code
// (compiled with -ffp-contract=off): coordinate-system axes from Euler angles and instance bounding
// boxes. Used by non-C++ hosts (Python) that build scenes without the reference's World.
//   CoordSystem::applyRotation / lookAt        /root/reference/RayZath/render_parts.cpp:51-62
//   vec3 rotations                              /root/reference/RayZath/cuda_render_parts.cuh:116-182 (sign conventions)
//   Instance::calculateBoundingBox              /root/reference/RayZath/instance.cpp:118-155
#include "../../include/rzb200.h"

#include <cmath>

namespace
{
	struct V3 { float x, y, z; };
	inline void rotX(V3& v, float a) { const float s = sinf(a), c = cosf(a); const float ny = v.y * c + v.z * s; v.z = v.y * -s + v.z * c; v.y = ny; }
	inline void rotY(V3& v, float a) { const float s = sinf(a), c = cosf(a); const float nx = v.x * c - v.z * s; v.z = v.x * s + v.z * c; v.x = nx; }
	inline void rotZ(V3& v, float a) { const float s = sinf(a), c = cosf(a); const float nx = v.x * c + v.y * s; v.y = v.x * -s + v.y * c; v.x = nx; }
}

// order: 0 = RotatedXYZ (instances, CoordSystem::applyRotation), 1 = Z then X then Y (cameras, CoordSystem::lookAt)
extern "C" int rzb_rotation_axes(const float* rotation, int order, float* axes_out)
{
	if (!rotation || !axes_out) return RZB_ERR_INVALID;
	for (int k = 0; k < 3; ++k)
	{
		V3 v{k == 0 ? 1.0f : 0.0f, k == 1 ? 1.0f : 0.0f, k == 2 ? 1.0f : 0.0f};
		if (order == 0) { rotX(v, rotation[0]); rotY(v, rotation[1]); rotZ(v, rotation[2]); }
		else { rotZ(v, rotation[2]); rotX(v, rotation[0]); rotY(v, rotation[1]); }
		axes_out[3 * k] = v.x; axes_out[3 * k + 1] = v.y; axes_out[3 * k + 2] = v.z;
	}
	return RZB_OK;
}

extern "C" int rzb_instance_bbox(const float* vertices, uint32_t nv, const float* position, const float* scale,
	const float* axes, float* bbox_out)
{
	if (!vertices || !position || !scale || !axes || !bbox_out) return RZB_ERR_INVALID;
	for (int k = 0; k < 6; ++k) bbox_out[k] = 0.0f;
	if (nv == 0) return RZB_OK;
	float mn[3], mx[3];
	for (uint32_t i = 0; i < nv; ++i)
	{
		const float vx = vertices[3 * size_t(i)] * scale[0];
		const float vy = vertices[3 * size_t(i) + 1] * scale[1];
		const float vz = vertices[3 * size_t(i) + 2] * scale[2];
		// x_axis * v.x + y_axis * v.y + z_axis * v.z
		float p[3];
		for (int k = 0; k < 3; ++k) p[k] = axes[k] * vx + axes[3 + k] * vy + axes[6 + k] * vz;
		if (i == 0) { for (int k = 0; k < 3; ++k) mn[k] = mx[k] = p[k]; continue; }
		for (int k = 0; k < 3; ++k)
		{
			if (mn[k] > p[k]) mn[k] = p[k];
			if (mx[k] < p[k]) mx[k] = p[k];
		}
	}
	for (int k = 0; k < 3; ++k) { bbox_out[k] = mn[k] + position[k]; bbox_out[3 + k] = mx[k] + position[k]; }
	return RZB_OK;
}

// Triangle::calculateNormal (mesh_component.cpp:19-26): cross(v2 - v3, v2 - v1), normalised by division
extern "C" int rzb_face_normals(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt, float* normals_out)
{
	if (!vertices || !tris || !normals_out) return RZB_ERR_INVALID;
	for (uint32_t i = 0; i < nt; ++i)
	{
		const uint32_t a = tris[3 * size_t(i)], b = tris[3 * size_t(i) + 1], c = tris[3 * size_t(i) + 2];
		if (a >= nv || b >= nv || c >= nv) return RZB_ERR_INVALID;
		const float* v1 = vertices + 3 * size_t(a);
		const float* v2 = vertices + 3 * size_t(b);
		const float* v3 = vertices + 3 * size_t(c);
		const float ax = v2[0] - v3[0], ay = v2[1] - v3[1], az = v2[2] - v3[2];
		const float bx = v2[0] - v1[0], by = v2[1] - v1[1], bz = v2[2] - v1[2];
		float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
		const float m = sqrtf(nx * nx + ny * ny + nz * nz);
		normals_out[3 * size_t(i)] = nx / m;
		normals_out[3 * size_t(i) + 1] = ny / m;
		normals_out[3 * size_t(i) + 2] = nz / m;
	}
	return RZB_OK;
}

// ---- refit: new boxes for an existing tree after its vertices moved. The topology (children,
// triangle ranges, order) is kept; every leaf box becomes the exact fp32 min / max of its triangles' vertices -- what the
// builders compute (bvh_tree_node.hpp: BoundingBox::extendBy over the objects' boxes) -- and every inner box the union
// of its children's. Iterative post-order walk from the root; works for any layout rzb_set_scene accepts.
#include <algorithm>
#include <utility>
#include <vector>

extern "C" int rzb_refit_mesh_bvh(const float* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt,
	rzb_node* nodes, uint32_t node_count, const uint32_t* order)
{
	if (!vertices || !tris || !nodes || !order || node_count == 0) return RZB_ERR_INVALID;
	std::vector<std::pair<uint32_t, bool>> todo; // (node, children done)
	std::vector<uint8_t> seen(node_count, 0);
	todo.emplace_back(0u, false);
	seen[0] = 1;
	while (!todo.empty())
	{
		const uint32_t i = todo.back().first;
		const bool children_done = todo.back().second;
		rzb_node& n = nodes[i];
		const uint32_t count = n.type_count & 0x3FFFFFFFu;
		if (count != 0u)
		{
			todo.pop_back();
			if (uint64_t(n.begin) + count > nt) return RZB_ERR_INVALID;
			float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
			for (uint32_t t = n.begin; t < n.begin + count; ++t)
			{
				const uint32_t tri = order[t];
				if (tri >= nt) return RZB_ERR_INVALID;
				for (int c = 0; c < 3; ++c)
				{
					const uint32_t v = tris[3 * size_t(tri) + c];
					if (v >= nv) return RZB_ERR_INVALID;
					for (int k = 0; k < 3; ++k)
					{
						mn[k] = std::min(mn[k], vertices[3 * size_t(v) + k]);
						mx[k] = std::max(mx[k], vertices[3 * size_t(v) + k]);
					}
				}
			}
			for (int k = 0; k < 3; ++k) { n.bb_min[k] = mn[k]; n.bb_max[k] = mx[k]; }
			continue;
		}
		if (uint64_t(n.begin) + 1 >= node_count) return RZB_ERR_INVALID;
		if (!children_done)
		{
			if (seen[n.begin] || seen[n.begin + 1]) return RZB_ERR_INVALID; // not a tree
			seen[n.begin] = seen[n.begin + 1] = 1;
			todo.back().second = true;
			todo.emplace_back(n.begin, false);
			todo.emplace_back(n.begin + 1u, false);
			continue;
		}
		todo.pop_back();
		const rzb_node& a = nodes[n.begin];
		const rzb_node& b = nodes[n.begin + 1];
		for (int k = 0; k < 3; ++k)
		{
			n.bb_min[k] = std::min(a.bb_min[k], b.bb_min[k]);
			n.bb_max[k] = std::max(a.bb_max[k], b.bb_max[k]);
		}
	}
	return RZB_OK;
}
