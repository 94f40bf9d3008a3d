// Host side of the wide (4-ary) own-tree mode (RZB_SCENE_WIDE_TREES): collapse of a binary mesh tree in rzb_node format
// into the 128-byte wide nodes rzb_traverse.cuh walks. Used by rzb_set_scene and by the host simulation of the traversal.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/rzb200.h"

namespace rzb
{
	constexpr uint32_t kWideLeafBitHost = 1u << 29;
	constexpr uint32_t kWideEmptyHost = 0x3FFFFFFFu;
	struct WideNode { float v[6][4]; uint32_t ref[4]; uint32_t pad[4]; };
	static_assert(sizeof(WideNode) == 128, "WideNode");
	// Collapse one binary mesh tree (rzb_node, local indices) into 4-ary nodes appended to `out`: a node's slots are its two
	// children, then -- while slots are free -- the inner slot with the largest box surface is replaced by ITS two children.
	// Returns the reference of binary node i (see rzb_traverse.cuh: wide_decode), depth through depth_out.
	inline uint32_t collapseWide(const rzb_node* nodes, const uint32_t i, const uint32_t tri_offset, std::vector<WideNode>& out,
		const uint32_t depth, uint32_t& depth_out, bool& ok)
	{
		const rzb_node& n = nodes[i];
		const uint32_t count = n.type_count & 0x3FFFFFFFu;
		if (count != 0u)
		{
			if (count > 15u || uint64_t(tri_offset) + n.begin + count > (1u << 25)) { ok = false; return kWideEmptyHost; }
			return kWideLeafBitHost | (count << 25) | (tri_offset + n.begin);
		}
		depth_out = std::max(depth_out, depth + 1u);
		uint32_t slots[4] = {n.begin, n.begin + 1u, 0u, 0u};
		int used = 2;
		auto area = [&](uint32_t k) {
			const rzb_node& c = nodes[k];
			const float dx = c.bb_max[0] - c.bb_min[0], dy = c.bb_max[1] - c.bb_min[1], dz = c.bb_max[2] - c.bb_min[2];
			return dx * dy + dy * dz + dz * dx;
		};
		while (used < 4)
		{
			int best = -1;
			float best_area = -1.0f;
			for (int k = 0; k < used; ++k)
				if ((nodes[slots[k]].type_count & 0x3FFFFFFFu) == 0u && area(slots[k]) > best_area) { best = k; best_area = area(slots[k]); }
			if (best < 0) break;
			const uint32_t b = nodes[slots[best]].begin;
			slots[best] = b;
			slots[used++] = b + 1u;
		}
		const size_t q = out.size();
		out.emplace_back();
		if (q >= (1u << 29) - 1u) { ok = false; return kWideEmptyHost; }
		WideNode w{};
		for (int k = 0; k < 4; ++k)
		{
			if (k < used)
			{
				const rzb_node& c = nodes[slots[k]];
				for (int a = 0; a < 3; ++a) { w.v[a][k] = c.bb_min[a]; w.v[3 + a][k] = c.bb_max[a]; }
				w.ref[k] = collapseWide(nodes, slots[k], tri_offset, out, depth + 1u, depth_out, ok);
			}
			else
			{
				for (int a = 0; a < 3; ++a) { w.v[a][k] = INFINITY; w.v[3 + a][k] = -INFINITY; }
				w.ref[k] = kWideEmptyHost;
			}
		}
		out[q] = w;
		return uint32_t(q);
	}
}
