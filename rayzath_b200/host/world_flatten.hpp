// Flattens a RayZath::Engine::World into the POD arrays of the C ABI (include/rzb200.h).
//
// Host side of the drop-in: this is what replaces the reference's device mirroring
//   World::reconstructResources / reconstructObjects   /root/reference/RayZath/cuda_world.cu
//   Mesh::reconstruct                                   /root/reference/RayZath/cuda_instance.cu:17-226
//   Instance::reconstruct                               /root/reference/RayZath/cuda_instance.cu:236-285
//   ObjectContainerWithBVH::reconstruct/constructNode   /root/reference/RayZath/cuda_bvh.cuh:30-111
//   Material/DirectLight/SpotLight/Camera::reconstruct  /root/reference/RayZath/cuda_*.cu
// It compiles against the reference's own headers where they lie (never copied) and walks the host
// BVHs the reference's World::update() built, so the device traverses the identical tree.
#ifndef RZB_WORLD_FLATTEN_HPP
#define RZB_WORLD_FLATTEN_HPP

#include "world.hpp"
#include "engine_parts.hpp"

#include "../../include/rzb200.h"

#include <cstring>
#include <unordered_map>
#include <vector>

namespace rzb_host
{
	namespace RZ = RayZath::Engine;

	struct FlatScene
	{
		std::vector<rzb_node> mesh_nodes;
		std::vector<rzb_triangle> triangles;
		std::vector<uint32_t> tri_host_index;
		std::vector<rzb_mesh> meshes;
		std::vector<rzb_node> instance_nodes;
		std::vector<rzb_instance> instances;
		std::vector<uint32_t> instance_materials;
		std::vector<rzb_material> materials;
		std::vector<rzb_map> maps;
		std::vector<rzb_direct_light> direct_lights;
		std::vector<rzb_spot_light> spot_lights;
		rzb_material world_material{};
		uint32_t default_material = 0;
		uint32_t scene_flags = RZB_SCENE_REFERENCE_TREES;

		rzb_scene view() const
		{
			rzb_scene s{};
			s.mesh_nodes = mesh_nodes.data(); s.mesh_node_count = uint32_t(mesh_nodes.size());
			s.triangles = triangles.data(); s.triangle_count = uint32_t(triangles.size());
			s.tri_host_index = tri_host_index.data();
			s.meshes = meshes.data(); s.mesh_count = uint32_t(meshes.size());
			s.instance_nodes = instance_nodes.data(); s.instance_node_count = uint32_t(instance_nodes.size());
			s.instances = instances.data(); s.instance_count = uint32_t(instances.size());
			s.instance_materials = instance_materials.data();
			s.instance_material_count = uint32_t(instance_materials.size());
			s.materials = materials.data(); s.material_count = uint32_t(materials.size());
			s.maps = maps.data(); s.map_count = uint32_t(maps.size());
			s.direct_lights = direct_lights.data(); s.direct_light_count = uint32_t(direct_lights.size());
			s.spot_lights = spot_lights.data(); s.spot_light_count = uint32_t(spot_lights.size());
			s.world_material = world_material;
			s.default_material = default_material;
			s.flags = scene_flags;
			return s;
		}
	};

	inline void put3(float* dst, const Math::vec3f& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

	inline rzb_node makeNode(const RZ::BoundingBox& bb, uint32_t type, uint32_t begin, uint32_t count)
	{
		rzb_node n;
		put3(n.bb_min, bb.min);
		put3(n.bb_max, bb.max);
		n.begin = begin;
		n.type_count = ((type << 30) & 0xC0000000u) | (count & 0x3FFFFFFFu);
		return n;
	}

	class WorldFlattener
	{
		const RZ::World& world;
		FlatScene& out;
		// per map container: host idx -> index into out.maps
		std::vector<uint32_t> tex_ids, nrm_ids, met_ids, rgh_ids, emi_ids;

		template <RZ::ObjectType T, typename Texel>
		void flattenMaps(std::vector<uint32_t>& ids, uint32_t format)
		{
			const auto& container = world.container<T>();
			ids.assign(container.count(), RZB_NO_INDEX);
			for (uint32_t i = 0; i < container.count(); ++i)
			{
				const auto& h = container[i];
				if (!h) continue;
				rzb_map m{};
				m.format = format;
				m.width = uint32_t(h->bitmap().GetWidth());
				m.height = uint32_t(h->bitmap().GetHeight());
				m.filter = h->filterMode() == RZ::TextureBufferBase::FilterMode::Linear ? RZB_FILTER_LINEAR : RZB_FILTER_POINT;
				switch (h->addressMode())
				{
					case RZ::TextureBufferBase::AddressMode::Wrap: m.address = RZB_ADDRESS_WRAP; break;
					case RZ::TextureBufferBase::AddressMode::Clamp: m.address = RZB_ADDRESS_CLAMP; break;
					case RZ::TextureBufferBase::AddressMode::Mirror: m.address = RZB_ADDRESS_MIRROR; break;
					default: m.address = RZB_ADDRESS_BORDER; break;
				}
				m.scale[0] = h->scale().x; m.scale[1] = h->scale().y;
				m.rotation = h->rotation().value();
				m.translation[0] = h->translation().x; m.translation[1] = h->translation().y;
				m.pixels = h->bitmap().GetMapAddress();
				static_assert(sizeof(Texel) == 1 || sizeof(Texel) == 4, "texel size");
				ids[i] = uint32_t(out.maps.size());
				out.maps.push_back(m);
			}
		}
		template <typename HandleT>
		static uint32_t mapId(const HandleT& h, const std::vector<uint32_t>& ids)
		{
			if (!h) return RZB_NO_INDEX;
			const uint32_t idx = h.accessor()->idx();
			return idx < ids.size() ? ids[idx] : RZB_NO_INDEX;
		}
		rzb_material flattenMaterial(const RZ::Material& m) const
		{
			rzb_material r{};
			const Graphics::Color c = m.color();
			r.color[0] = c.red / 255.0f; r.color[1] = c.green / 255.0f;
			r.color[2] = c.blue / 255.0f; r.color[3] = c.alpha / 255.0f;
			r.metalness = m.metalness(); r.roughness = m.roughness(); r.emission = m.emission();
			r.ior = m.ior(); r.scattering = m.scattering();
			r.texture = mapId(m.map<RZ::ObjectType::Texture>(), tex_ids);
			r.normal_map = mapId(m.map<RZ::ObjectType::NormalMap>(), nrm_ids);
			r.metalness_map = mapId(m.map<RZ::ObjectType::MetalnessMap>(), met_ids);
			r.roughness_map = mapId(m.map<RZ::ObjectType::RoughnessMap>(), rgh_ids);
			r.emission_map = mapId(m.map<RZ::ObjectType::EmissionMap>(), emi_ids);
			return r;
		}

		// ---- meshes: Mesh::reconstruct order (root, then sibling pairs, subtree(first), subtree(second)) ----
		using tri_node_t = RZ::ComponentTreeNode<RZ::Triangle>;
		void addTriangle(const RZ::Mesh& mesh, const RZ::Triangle& t)
		{
			rzb_triangle d{};
			if (t.areVertsValid())
				for (int k = 0; k < 3; ++k) put3(d.v[k], mesh.vertices()[t.vertices[k]]);
			else
			{
				d.v[1][0] = 1.0f; d.v[2][1] = 1.0f;
			}
			if (t.areTexcrdsValid())
				for (int k = 0; k < 3; ++k)
				{
					d.uv[k][0] = mesh.texcrds()[t.texcrds[k]].x;
					d.uv[k][1] = mesh.texcrds()[t.texcrds[k]].y;
				}
			else
			{
				d.uv[1][1] = 1.0f; d.uv[2][0] = 1.0f;
			}
			if (t.areNormalsValid())
				for (int k = 0; k < 3; ++k) put3(d.n[k], mesh.normals()[t.normals[k]]);
			else
				for (int k = 0; k < 3; ++k) put3(d.n[k], t.normal);
			put3(d.face_normal, t.normal);
			d.material_slot = t.material_id & 0x3Fu; // cuda_render_parts.cu:8-12
			out.triangles.push_back(d);
			out.tri_host_index.push_back(uint32_t(&t - &mesh.triangles()[0]));
		}
		void addLeaf(const RZ::Mesh& mesh, const tri_node_t& n, uint32_t node_base, uint32_t tri_base)
		{
			(void)node_base;
			// the leaf's count is what is actually emitted: a null component pointer in the leaf is skipped (as the
			// instance path below does), so the range never reaches into the next leaf's triangles
			const uint32_t begin = uint32_t(out.triangles.size());
			const size_t node_at = out.mesh_nodes.size();
			out.mesh_nodes.push_back(makeNode(n.boundingBox(), 0, begin - tri_base, 0));
			for (const auto* object : n.objects())
				if (object) addTriangle(mesh, *object);
			out.mesh_nodes[node_at].type_count = uint32_t(out.triangles.size()) - begin;
		}
		void buildChildren(const RZ::Mesh& mesh, const tri_node_t& n, uint32_t node_base, uint32_t tri_base)
		{
			const auto& c1 = n.children()->first;
			const uint32_t first_subtree = c1.treeSize() - 1;
			if (c1.isLeaf()) addLeaf(mesh, c1, node_base, tri_base);
			else out.mesh_nodes.push_back(makeNode(c1.boundingBox(), uint32_t(c1.children()->type),
				uint32_t(out.mesh_nodes.size()) - node_base + 2, 0));
			const auto& c2 = n.children()->second;
			if (c2.isLeaf()) addLeaf(mesh, c2, node_base, tri_base);
			else out.mesh_nodes.push_back(makeNode(c2.boundingBox(), uint32_t(c2.children()->type),
				uint32_t(out.mesh_nodes.size()) - node_base + first_subtree + 1, 0));
			if (!c1.isLeaf()) buildChildren(mesh, c1, node_base, tri_base);
			if (!c2.isLeaf()) buildChildren(mesh, c2, node_base, tri_base);
		}
		// Optional: ignore the host tree and build this repo's SAH tree over the same triangles (rzb_build_mesh_bvh_sah;
		// RZB_SCENE_OWN_TREES). The triangle records are the same, only their order and the nodes differ.
		// Compiled only where librzb200.so is linked (the drop-in build defines RZB_FLATTEN_WITH_OWN_TREES); the oracle's
		// rz_ref_tool includes this header for the reference trees alone and must not depend on the product library.
		bool flattenMeshOwnTree(const RZ::Mesh& mesh, rzb_mesh& m)
		{
#ifndef RZB_FLATTEN_WITH_OWN_TREES
			(void)mesh; (void)m;
			return false;
#else
			const uint32_t nt = mesh.triangles().count(), nv = mesh.vertices().count();
			std::vector<float> verts(size_t(nv) * 3);
			for (uint32_t i = 0; i < nv; ++i) put3(&verts[size_t(i) * 3], mesh.vertices()[i]);
			std::vector<uint32_t> ids(size_t(nt) * 3);
			for (uint32_t i = 0; i < nt; ++i)
			{
				const RZ::Triangle& t = mesh.triangles()[i];
				if (!t.areVertsValid()) return false; // keep the host tree for meshes with placeholder triangles
				for (int k = 0; k < 3; ++k) ids[size_t(i) * 3 + k] = t.vertices[k];
			}
			std::vector<rzb_node> nodes(size_t(nt) * 2 + 1);
			std::vector<uint32_t> order(nt);
			uint32_t count = 0;
			if (rzb_build_mesh_bvh_sah(verts.data(), nv, ids.data(), nt, 4, nodes.data(), uint32_t(nodes.size()), &count, order.data()) != RZB_OK)
				return false;
			out.mesh_nodes.insert(out.mesh_nodes.end(), nodes.begin(), nodes.begin() + count);
			for (uint32_t i = 0; i < nt; ++i) addTriangle(mesh, mesh.triangles()[order[i]]);
			return true;
#endif
		}
		void flattenMesh(const RZ::Mesh& mesh)
		{
			rzb_mesh m{};
			m.node_offset = uint32_t(out.mesh_nodes.size());
			m.tri_offset = uint32_t(out.triangles.size());
			const auto& root = mesh.triangles().getBVH().rootNode();
			if (own_trees && mesh.triangles().count() != 0 && flattenMeshOwnTree(mesh, m))
			{
				out.scene_flags = RZB_SCENE_OWN_TREES;
			}
			else if (mesh.triangles().count() != 0)
			{
				if (root.isLeaf()) addLeaf(mesh, root, m.node_offset, m.tri_offset);
				else
				{
					out.mesh_nodes.push_back(makeNode(root.boundingBox(), uint32_t(root.children()->type), 1, 0));
					buildChildren(mesh, root, m.node_offset, m.tri_offset);
				}
			}
			m.node_count = uint32_t(out.mesh_nodes.size()) - m.node_offset;
			m.tri_count = uint32_t(out.triangles.size()) - m.tri_offset;
			out.meshes.push_back(m);
		}

		// ---- instances: ObjectContainerWithBVH::constructNode order ----
		using inst_node_t = RZ::TreeNode<RZ::Instance>;
		void addInstance(const RZ::Handle<RZ::Instance>& h, const std::vector<uint32_t>& mesh_ids,
			const std::vector<uint32_t>& material_ids)
		{
			rzb_instance d{};
			const RZ::Transformation& t = h->transformationInGroup();
			put3(d.position, t.position());
			put3(d.scale, t.scale());
			put3(d.axis_x, t.coordSystem().xAxis());
			put3(d.axis_y, t.coordSystem().yAxis());
			put3(d.axis_z, t.coordSystem().zAxis());
			put3(d.bb_min, h->boundingBox().min);
			put3(d.bb_max, h->boundingBox().max);
			d.mesh = RZB_NO_INDEX;
			if (h->mesh())
			{
				const uint32_t idx = h->mesh().accessor()->idx();
				if (idx < mesh_ids.size()) d.mesh = mesh_ids[idx];
			}
			d.material_offset = uint32_t(out.instance_materials.size());
			uint32_t used = 0;
			for (uint32_t i = 0; i < RZ::Instance::materialCapacity(); ++i)
				if (h->material(i)) used = i + 1;
			for (uint32_t i = 0; i < used; ++i)
			{
				uint32_t id = out.default_material;
				if (const auto& mat = h->material(i); mat)
				{
					const uint32_t idx = mat.accessor()->idx();
					if (idx < material_ids.size() && material_ids[idx] != RZB_NO_INDEX) id = material_ids[idx];
				}
				out.instance_materials.push_back(id);
			}
			d.material_count = used;
			d.host_index = h.accessor()->idx();
			out.instances.push_back(d);
		}
		void constructInstanceNode(size_t slot, const inst_node_t& n, const std::vector<uint32_t>& mesh_ids,
			const std::vector<uint32_t>& material_ids)
		{
			if (n.isLeaf())
			{
				uint32_t count = 0;
				const uint32_t begin = uint32_t(out.instances.size());
				for (const auto& object : n.objects())
				{
					// the reference counts every handle of the leaf (cuda_bvh.cuh:95-99); dead handles
					// cannot be mirrored, so they are dropped here and the count follows.
					if (!object) continue;
					addInstance(object, mesh_ids, material_ids);
					++count;
				}
				out.instance_nodes[slot] = makeNode(n.boundingBox(), 0, begin, count);
			}
			else
			{
				const size_t first = out.instance_nodes.size();
				out.instance_nodes[slot] = makeNode(n.boundingBox(), uint32_t(n.children()->type), uint32_t(first), 0);
				out.instance_nodes.emplace_back();
				out.instance_nodes.emplace_back();
				constructInstanceNode(first, n.children()->first, mesh_ids, material_ids);
				constructInstanceNode(first + 1, n.children()->second, mesh_ids, material_ids);
			}
		}

	public:
		// own_trees: build this repo's SAH triangle trees instead of flattening the host's (the reference's) trees
		bool own_trees = false;
		WorldFlattener(const RZ::World& w, FlatScene& o, bool own_trees_ = false) : world(w), out(o), own_trees(own_trees_) {}

		// Precondition: world.update() has been called (host BVHs are current), cuda_engine_core.cu:58-60.
		// keep_geometry: the Mesh container has not changed since the previous run into the same FlatScene -- its
		// mesh_nodes / triangles / tri_host_index / meshes stay as they are (RZB_SCENE_KEEP_GEOMETRY), everything else is
		// flattened again.
		void run(const bool keep_geometry = false)
		{
			if (keep_geometry)
			{
				FlatScene kept;
				kept.mesh_nodes.swap(out.mesh_nodes);
				kept.triangles.swap(out.triangles);
				kept.tri_host_index.swap(out.tri_host_index);
				kept.meshes.swap(out.meshes);
				kept.scene_flags = out.scene_flags;
				out = std::move(kept);
			}
			else out = FlatScene{};
			flattenMaps<RZ::ObjectType::Texture, Graphics::Color>(tex_ids, RZB_MAP_RGBA8);
			flattenMaps<RZ::ObjectType::NormalMap, Graphics::Color>(nrm_ids, RZB_MAP_RGBA8);
			flattenMaps<RZ::ObjectType::MetalnessMap, uint8_t>(met_ids, RZB_MAP_R8);
			flattenMaps<RZ::ObjectType::RoughnessMap, uint8_t>(rgh_ids, RZB_MAP_R8);
			flattenMaps<RZ::ObjectType::EmissionMap, float>(emi_ids, RZB_MAP_R32F);

			const auto& materials = world.container<RZ::ObjectType::Material>();
			std::vector<uint32_t> material_ids(materials.count(), RZB_NO_INDEX);
			for (uint32_t i = 0; i < materials.count(); ++i)
			{
				if (!materials[i]) continue;
				material_ids[i] = uint32_t(out.materials.size());
				out.materials.push_back(flattenMaterial(*materials[i]));
			}
			out.default_material = uint32_t(out.materials.size());
			out.materials.push_back(flattenMaterial(world.defaultMaterial()));
			out.world_material = flattenMaterial(world.material());

			const auto& meshes = world.container<RZ::ObjectType::Mesh>();
			std::vector<uint32_t> mesh_ids(meshes.count(), RZB_NO_INDEX);
			uint32_t next_mesh = 0;
			for (uint32_t i = 0; i < meshes.count(); ++i)
			{
				if (!meshes[i]) continue;
				mesh_ids[i] = next_mesh++;
				if (!keep_geometry) flattenMesh(*meshes[i]);
			}

			const auto& dls = world.container<RZ::ObjectType::DirectLight>();
			for (uint32_t i = 0; i < dls.count(); ++i)
			{
				if (!dls[i]) continue;
				rzb_direct_light d{};
				put3(d.direction, dls[i]->direction());
				d.angular_size = dls[i]->angularSize();
				const Graphics::Color c = dls[i]->color();
				d.color[0] = c.red / 255.0f; d.color[1] = c.green / 255.0f; d.color[2] = c.blue / 255.0f;
				d.emission = dls[i]->emission();
				out.direct_lights.push_back(d);
			}
			const auto& sls = world.container<RZ::ObjectType::SpotLight>();
			for (uint32_t i = 0; i < sls.count(); ++i)
			{
				if (!sls[i]) continue;
				rzb_spot_light d{};
				put3(d.position, sls[i]->position());
				put3(d.direction, sls[i]->direction());
				d.size = sls[i]->size();
				d.beam_angle = sls[i]->GetBeamAngle();
				const Graphics::Color c = sls[i]->color();
				d.color[0] = c.red / 255.0f; d.color[1] = c.green / 255.0f; d.color[2] = c.blue / 255.0f;
				d.emission = sls[i]->emission();
				out.spot_lights.push_back(d);
			}

			const auto& instances = world.container<RZ::ObjectType::Instance>();
			if (instances.count() != 0)
			{
				out.instance_nodes.emplace_back();
				constructInstanceNode(0, instances.root(), mesh_ids, material_ids);
			}
		}
	};

	inline rzb_camera flattenCamera(const RZ::Camera& c)
	{
		rzb_camera r{};
		r.width = c.width(); r.height = c.height();
		put3(r.position, c.position());
		put3(r.axis_x, c.coordSystem().xAxis());
		put3(r.axis_y, c.coordSystem().yAxis());
		put3(r.axis_z, c.coordSystem().zAxis());
		r.fov = c.fov().value();
		r.near_far[0] = c.nearFar().x; r.near_far[1] = c.nearFar().y;
		r.focal_distance = c.focalDistance();
		r.aperture = c.aperture();
		r.exposure_time = c.exposureTime();
		r.temporal_blend = c.temporalBlend();
		r.raycast_pixel[0] = c.getRayCastPixel().x; r.raycast_pixel[1] = c.getRayCastPixel().y;
		return r;
	}

	inline rzb_config flattenConfig(const RZ::RenderConfig& rc, uint64_t seed)
	{
		rzb_config c{};
		c.spot_light_samples = rc.lightSampling().spotLight();
		c.direct_light_samples = rc.lightSampling().directLight();
		c.max_depth = rc.tracing().maxDepth();
		c.flags = RZB_FLAG_NONE;
		c.seed = seed;
		return c;
	}
}
#endif
