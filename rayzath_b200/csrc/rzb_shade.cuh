// Shading half of the wavefront path tracer: RNG, map fetches, the uber-material, BSDF sampling,
// next-event estimation with MIS, camera ray generation, any-hit shadow traversal.
//
// Follows the integrator of the reference's CUDA engine (the path being replaced):
//   traceRay / directIllumination / *LightSampling   /root/reference/RayZath/cuda_render_kernel.cu:146-355
//   Material (opacityColor, BRDF, NDF, samplers)     /root/reference/RayZath/cuda_material.cuh:70-301
//   sampling helpers, fresnel                        /root/reference/RayZath/cuda_render_parts.cuh:1196-1353
//   lights                                           /root/reference/RayZath/cuda_direct_light.cuh:50-74, cuda_spot_light.cuh:56-86
//   camera rays                                      /root/reference/RayZath/cuda_camera.cuh:303-379
//   surface analysis                                 /root/reference/RayZath/cuda_instance.cuh:231-263
// With RZB_FLAG_CPU_SEMANTICS the places where cpu_engine_kernel.cpp differs are followed instead
// (SURVEY.md §8a "divergences"): no medium scattering, no Beer-Lambert, opaque shadows, texture and
// emission maps replace instead of multiply.
//
// Deliberate differences: the RNG is a counter-based hash keyed on
// (seed, pixel, pass, dimension) instead of the reference's unseeded 2-float multiplicative generator, so
// renders are reproducible and sample streams of different GPUs are disjoint by construction.
#pragma once

#include "rzb_traverse.cuh"

namespace rzb
{
	// ------------------------------------------------------------------ small float3 helpers (fast path, FMA allowed)
	__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
	__device__ __forceinline__ float3 f3(const V3& v) { return make_float3(v.x, v.y, v.z); }
	__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
	__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
	__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
	__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
	__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
	__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
	__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
	__device__ __forceinline__ float3 cross(float3 a, float3 b)
	{
		return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
	}
	__device__ __forceinline__ float length(float3 a) { return sqrtf(dot(a, a)); }
	__device__ __forceinline__ float3 normalize(float3 a) { return a * rsqrtf(dot(a, a)); }
	__device__ __forceinline__ float similarity(float3 a, float3 b) { return dot(a, b) * rsqrtf(dot(a, a)) * rsqrtf(dot(b, b)); }
	__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + (b - a) * t; }
	__device__ __forceinline__ float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }

	// ------------------------------------------------------------------ RNG
	__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
	{
		x ^= x >> 16; x *= 0x7feb352du;
		x ^= x >> 15; x *= 0x846ca68bu;
		x ^= x >> 16;
		return x;
	}
	struct Rng
	{
		uint32_t base, dim;
		__host__ __device__ __forceinline__ Rng(uint64_t seed, uint32_t pixel, uint32_t pass)
		{
			uint32_t h = mix32(uint32_t(seed) ^ (pixel * 0x9E3779B1u));
			h = mix32(h ^ (pass * 0x85EBCA77u) ^ uint32_t(seed >> 32));
			base = h;
			dim = 0u;
		}
		// uniform in [0, 1)
		__host__ __device__ __forceinline__ float next()
		{
			const uint32_t x = mix32(base ^ (++dim * 0xC2B2AE3Du));
			return float(x >> 8) * (1.0f / 16777216.0f);
		}
		__host__ __device__ __forceinline__ float next_signed() { return next() * 2.0f - 1.0f; }
	};

	// ------------------------------------------------------------------ maps (render_parts.hpp:209-221, cuda_buffer.cuh:427-438)
	__device__ __forceinline__ int address_texel(int i, const int n, const uint32_t mode, bool& inside)
	{
		inside = true;
		switch (mode)
		{
			case RZB_ADDRESS_WRAP:
				if (unsigned(i) < unsigned(n)) return i; // the common case without the emulated integer division
				i %= n;
				return i < 0 ? i + n : i;
			case RZB_ADDRESS_MIRROR:
			{
				int p = i % (2 * n);
				if (p < 0) p += 2 * n;
				return p < n ? p : 2 * n - 1 - p;
			}
			case RZB_ADDRESS_CLAMP:
				return min(max(i, 0), n - 1);
			default: // border
				inside = i >= 0 && i < n;
				return min(max(i, 0), n - 1);
		}
	}
	__device__ __forceinline__ float4 load_texel(const DMap& m, int x, int y)
	{
		bool in_x, in_y;
		x = address_texel(x, int(m.width), m.address, in_x);
		y = address_texel(y, int(m.height), m.address, in_y);
		if (!(in_x && in_y)) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		const size_t i = size_t(y) * m.width + size_t(x);
		if (m.format == RZB_MAP_RGBA8)
		{
			const uchar4 c = __ldg(reinterpret_cast<const uchar4*>(m.pixels) + i);
			return make_float4(c.x * (1.0f / 255.0f), c.y * (1.0f / 255.0f), c.z * (1.0f / 255.0f), c.w * (1.0f / 255.0f));
		}
		if (m.format == RZB_MAP_R8)
		{
			const float v = __ldg(reinterpret_cast<const unsigned char*>(m.pixels) + i) * (1.0f / 255.0f);
			return make_float4(v, v, v, v);
		}
		const float v = __ldg(reinterpret_cast<const float*>(m.pixels) + i);
		return make_float4(v, v, v, v);
	}
	__device__ __forceinline__ float4 fetch_map(const DMap& m, float u, float v)
	{
		u += m.trans_x; v += m.trans_y;
		const float ru = u * m.rot_cos + v * m.rot_sin;
		const float rv = v * m.rot_cos - u * m.rot_sin;
		u = ru * m.scale_x; v = rv * m.scale_y;
		if (m.filter == RZB_FILTER_POINT && m.address == RZB_ADDRESS_WRAP)
		{
			// the CPU engine's only mode (render_parts.hpp:215-220)
			// fmodf(a, 1) == a - truncf(a) exactly (the difference is representable): same values as the CPU engine's
			// fmodf chain without libdevice's iterative fmodf (ncu: 9 % of k_shade's instructions)
			const float fu = u - truncf(u) + 1.0f, fv = v - truncf(v) + 1.0f;
			const float x = fu - truncf(fu);
			const float y = 1.0f - (fv - truncf(fv));
			const int px = min(int(x * float(m.width)), int(m.width) - 1);
			const int py = min(int(y * float(m.height)), int(m.height) - 1);
			return load_texel(m, max(px, 0), max(py, 0));
		}
		v = 1.0f - v; // tex2D(u, 1 - v)
		if (m.address == RZB_ADDRESS_CLAMP)
		{
			u = fminf(fmaxf(u, 0.0f), 1.0f);
			v = fminf(fmaxf(v, 0.0f), 1.0f);
		}
		if (m.filter == RZB_FILTER_POINT)
			return load_texel(m, int(floorf(u * float(m.width))), int(floorf(v * float(m.height))));
		const float xb = u * float(m.width) - 0.5f, yb = v * float(m.height) - 0.5f;
		const float xf = floorf(xb), yf = floorf(yb);
		const float a = xb - xf, b = yb - yf;
		const int x0 = int(xf), y0 = int(yf);
		const float4 t00 = load_texel(m, x0, y0), t10 = load_texel(m, x0 + 1, y0);
		const float4 t01 = load_texel(m, x0, y0 + 1), t11 = load_texel(m, x0 + 1, y0 + 1);
		const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
		return make_float4(
			t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11,
			t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11,
			t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11,
			t00.w * w00 + t10.w * w10 + t01.w * w01 + t11.w * w11);
	}

	// ------------------------------------------------------------------ material (cuda_material.cuh:70-123, cpu_engine_kernel.cpp:505-537)
	__device__ __forceinline__ float4 material_opacity_color(const DScene& sc, const rzb_material& m, float u, float v)
	{
		float4 c = make_float4(m.color[0], m.color[1], m.color[2], 1.0f - m.color[3]);
		if (m.texture != kNoIndex)
		{
			float4 t = fetch_map(sc.maps[m.texture], u, v);
			t.w = 1.0f - t.w;
			if (sc.flags & RZB_FLAG_CPU_SEMANTICS) c = t;
			else c = make_float4(c.x * t.x, c.y * t.y, c.z * t.z, c.w * t.w);
		}
		return c;
	}
	__device__ __forceinline__ float material_emission(const DScene& sc, const rzb_material& m, float u, float v)
	{
		float e = m.emission;
		if (m.emission_map != kNoIndex)
		{
			const float t = fetch_map(sc.maps[m.emission_map], u, v).x;
			e = (sc.flags & RZB_FLAG_CPU_SEMANTICS) ? t : e * t;
		}
		return e;
	}
	__device__ __forceinline__ float material_metalness(const DScene& sc, const rzb_material& m, float u, float v)
	{
		return m.metalness_map != kNoIndex ? fetch_map(sc.maps[m.metalness_map], u, v).x : m.metalness;
	}
	__device__ __forceinline__ float material_roughness(const DScene& sc, const rzb_material& m, float u, float v)
	{
		return m.roughness_map != kNoIndex ? fetch_map(sc.maps[m.roughness_map], u, v).x : m.roughness;
	}
	__device__ __forceinline__ uint32_t instance_material(const DScene& sc, const uint32_t mat_offset,
		const uint32_t mat_count, const uint32_t slot)
	{
		return slot < mat_count ? __ldg(sc.inst_materials + mat_offset + slot) : sc.default_material;
	}

	struct Surface
	{
		uint32_t surface_material, behind_material;
		float u, v;
		float3 normal, mapped_normal;
		float3 color; float color_alpha;
		float metalness, roughness, emission;
		float fresnel, reflectance, tint_factor;
		float refr_x, refr_y;
	};

	// ------------------------------------------------------------------ sampling helpers (cuda_render_parts.cuh:1196-1353)
	__device__ __forceinline__ float3 reflect_vector(float3 vI, float3 vN) { return vN * (-2.0f * dot(vN, vI)) + vI; }
	__device__ __forceinline__ void local_coordinate(float3 vN, float3& vX, float3& vY)
	{
		const bool b = fabsf(vN.x) > fabsf(vN.y);
		vX = f3(b ? 0.0f : 1.0f, b ? 1.0f : 0.0f, 0.0f);
		vY = cross(vN, vX);
		vX = cross(vN, vY);
	}
	__device__ __forceinline__ float3 cosine_sample_hemisphere(float r1, float r2, float3 vN)
	{
		float3 vX, vY;
		local_coordinate(vN, vX, vY);
		float s, c;
		__sincosf(r1 * 6.283185f, &s, &c);
		const float sq = sqrtf(r2);
		return vX * (sq * c) + vY * (sq * s) + vN * sqrtf(1.0f - r2);
	}
	__device__ __forceinline__ float3 sample_sphere(float r1, float r2, float3 vN)
	{
		float3 vX, vY;
		local_coordinate(vN, vX, vY);
		float sp, cp;
		__sincosf(r1 * 6.283185f, &sp, &cp);
		// theta = acos(1 - 2 r2): cos(theta) = 1 - 2 r2, sin(theta) = sqrt(1 - cos^2) for theta in [0, pi]
		const float ct = 1.0f - 2.0f * r2;
		const float st = sqrtf(fmaxf(0.0f, 1.0f - ct * ct));
		return vX * (st * cp) + vY * (st * sp) + vN * ct;
	}
	__device__ __forceinline__ float3 sample_hemisphere(float r1, float r2, float3 vN) { return sample_sphere(r1, r2 * 0.5f, vN); }
	__device__ __forceinline__ float3 sample_disk(float3 vN, float radius, Rng& rng)
	{
		float3 vX, vY;
		local_coordinate(vN, vX, vY);
		const float r1 = rng.next() * 6.2831853f;
		const float r2 = rng.next();
		float s, c;
		__sincosf(r1, &s, &c);
		return (vX * s + vY * c) * (sqrtf(r2) * radius);
	}
	__device__ __forceinline__ float fresnel_specular_ratio(float3 vN, float3 vI, float n1, float n2, float& fx, float& fy)
	{
		const float ratio = n1 / n2;
		const float cosi = fabsf(dot(vI, vN));
		const float sin2_t = ratio * ratio * (1.0f - cosi * cosi);
		if (sin2_t >= 1.0f) return 1.0f;
		const float cost = sqrtf(1.0f - sin2_t);
		const float Rp = ((n1 * cosi) - (n2 * cost)) / ((n1 * cosi) + (n2 * cost));
		const float Rs = ((n2 * cosi) - (n1 * cost)) / ((n2 * cosi) + (n1 * cost));
		fx = ratio;
		fy = ratio * cosi - cost;
		return (Rs * Rs + Rp * Rp) * 0.5f;
	}

	// ------------------------------------------------------------------ BRDF (cuda_material.cuh:162-198)
	__device__ __forceinline__ float brdf(const Surface& s, const float scattering, float3 ray_dir, float3 vPL)
	{
		if (scattering > 0.0f) return 1.0f;
		const float n_o = dot(s.mapped_normal, vPL);
		if (n_o <= 0.0f) return 0.0f;
		const float n_i = dot(s.mapped_normal, -ray_dir);
		if (n_i <= 0.0f) return 0.0f;
		const float3 h = normalize(vPL - ray_dir); // halfwayVector
		const float n_h = dot(s.mapped_normal, h);
		const float b = (n_h * n_h) * (s.roughness - 1.0f) + 1.0001f;
		const float ndf = (s.roughness + 1.0e-5f) / (b * b);
		const float att_i = n_i / ((n_i * (1.0f - s.roughness)) + s.roughness);
		const float att_o = n_o / ((n_o * (1.0f - s.roughness)) + s.roughness);
		const float diffuse = n_o * float(s.color_alpha == 0.0f);
		const float specular = ndf * (att_i * att_o) / (n_i * n_o);
		return lerpf(diffuse, specular * n_o, s.reflectance);
	}

	// ------------------------------------------------------------------ direction sampling (cuda_material.cuh:203-301)
	// returns next direction; may switch the ray's medium and flip the geometric normal
	__device__ __forceinline__ float3 sample_direction(const DScene& sc, Surface& s, const float3 ray_dir,
		uint32_t& ray_medium, Rng& rng)
	{
		const float scattering = sc.materials[s.surface_material].scattering;
		if (s.color_alpha > 0.0f)
		{
			if (scattering > 0.0f)
			{
				const float r1 = rng.next(), r2 = rng.next();
				s.tint_factor = s.metalness;
				return sample_sphere(r1, r2, ray_dir);
			}
			if (s.fresnel < rng.next())
			{
				const float3 vO = ray_dir * s.refr_x + s.mapped_normal * s.refr_y;
				ray_medium = s.behind_material;
				s.normal = -s.normal;
				s.tint_factor = 1.0f;
				return vO;
			}
			float3 vO = reflect_vector(ray_dir, s.mapped_normal);
			const float d = dot(vO, s.normal);
			if (d < 0.0f) vO = vO + s.normal * (-2.0f * d);
			s.tint_factor = s.metalness;
			return vO;
		}
		if (rng.next() > s.reflectance)
		{
			const float r1 = rng.next(), r2 = rng.next();
			float3 vO = cosine_sample_hemisphere(r1, r2, s.mapped_normal);
			const float d = similarity(vO, s.normal);
			if (d < 0.0f) vO = vO + s.normal * (-2.0f * d);
			s.tint_factor = 1.0f;
			return vO;
		}
		const float r1 = rng.next();
		const float r2 = 1.0f - __powf(rng.next() + 1.0e-5f, s.roughness);
		const float3 vH = sample_hemisphere(r1, r2, s.mapped_normal);
		float3 vO = reflect_vector(ray_dir, vH);
		const float d = similarity(vO, s.normal);
		if (d < 0.0f) vO = vO + s.normal * (-2.0f * d);
		s.tint_factor = s.metalness;
		return vO;
	}

	// ------------------------------------------------------------------ surface analysis (cuda_instance.cuh:231-263)
	__device__ __forceinline__ void analyze_intersection(const DScene& sc, const uint32_t inst_idx, const uint32_t tri,
		const float b1, const float b2, const bool external, Surface& s)
	{
		const DInstance in = load_instance(sc.instances, inst_idx);
		const float4 h0 = __ldg(sc.tri_hot + 3 * size_t(tri));
		const float4 h1 = __ldg(sc.tri_hot + 3 * size_t(tri) + 1);
		const float4 h2 = __ldg(sc.tri_hot + 3 * size_t(tri) + 2);
		const float4 c0 = __ldg(sc.tri_cold + 5 * size_t(tri));
		const float4 c1 = __ldg(sc.tri_cold + 5 * size_t(tri) + 1);
		const float4 c2 = __ldg(sc.tri_cold + 5 * size_t(tri) + 2);
		const float4 c3 = __ldg(sc.tri_cold + 5 * size_t(tri) + 3);
		const float4 c4 = __ldg(sc.tri_cold + 5 * size_t(tri) + 4);
		const uint32_t slot = __float_as_uint(h2.y);
		s.surface_material = instance_material(sc, in.mat_offset, in.mat_count, slot);
		if (external) s.behind_material = s.surface_material;
		else if (sc.flags & RZB_FLAG_CPU_SEMANTICS) s.behind_material = sc.world_material;

		const float b3 = 1.0f - b1 - b2;
		const float2 t1 = make_float2(c0.w, c1.w), t2 = make_float2(c2.w, c3.w), t3 = make_float2(c4.x, c4.y);
		s.u = t1.x * b3 + t2.x * b1 + t3.x * b2;
		s.v = t1.y * b3 + t2.y * b1 + t3.y * b2;

		const float ext = external ? 1.0f : -1.0f;
		const float3 scale = f3(in.sx, in.sy, in.sz);
		const float3 ax = f3(in.xx, in.xy, in.xz), ay = f3(in.yx, in.yy, in.yz), az = f3(in.zx, in.zy, in.zz);
		float3 n = normalize(f3(c0.x, c0.y, c0.z) * b3 + f3(c1.x, c1.y, c1.z) * b1 + f3(c2.x, c2.y, c2.z) * b2);
		const rzb_material& mat = sc.materials[s.surface_material];
		if (mat.normal_map != kNoIndex)
		{
			// Triangle::mapNormal (cuda_render_parts.cuh:1095-1116)
			const float4 mc = fetch_map(sc.maps[mat.normal_map], s.u, s.v);
			const float3 e1 = f3(h0.w, h1.x, h1.y) * scale;
			const float3 e2 = f3(h1.z, h1.w, h2.x) * scale;
			const float2 d1 = make_float2(t2.x - t1.x, t2.y - t1.y), d2 = make_float2(t3.x - t1.x, t3.y - t1.y);
			n = n / scale;
			const float f = 1.0f / (d1.x * d2.y - d2.x * d1.y);
			float3 tangent = normalize((e1 * d2.y - e2 * d1.y) * f);
			tangent = normalize(tangent - n * dot(tangent, n));
			const float3 bitangent = cross(tangent, n);
			const float3 mn = f3(mc.x, mc.y, mc.z) * 2.0f - f3(1.0f, 1.0f, 1.0f);
			n = n * mn.z + tangent * mn.x + bitangent * mn.y;
			n = ax * n.x + ay * n.y + az * n.z; // transformL2GNoScale
		}
		else
		{
			n = n / scale;
			n = ax * n.x + ay * n.y + az * n.z; // transformL2G
		}
		s.mapped_normal = normalize(n) * ext;
		float3 g = f3(c3.x, c3.y, c3.z) * ext;
		g = g / scale;
		g = ax * g.x + ay * g.y + az * g.z;
		s.normal = normalize(g);
	}

	// ------------------------------------------------------------------ camera (cuda_camera.cuh:303-379, cpu_engine_kernel.cpp:180-252)
	struct DCamera
	{
		uint32_t width, height;
		float px, py, pz;
		float xx, xy, xz, yx, yy, yz, zx, zy, zz;
		float tana, aspect;
		float near_, far_;
		float focal_distance, aperture;
	};
	// pixel-centre ray; IEEE ops in the CPU engine's order so the primary-ray set can be compared bitwise
	__device__ __forceinline__ void camera_simple_ray(const DCamera& c, uint32_t x, uint32_t y, V3& o, V3& d)
	{
		const float dx = fmul(fsub(fdiv(fadd(float(x), 0.5f), float(c.width)), 0.5f), c.tana);
		const float dy = fmul(fsub(fdiv(fadd(float(y), 0.5f), float(c.height)), 0.5f), fdiv(-c.tana, c.aspect));
		V3 r = v3(
			fadd(fadd(fmul(c.xx, dx), fmul(c.yx, dy)), fmul(c.zx, 1.0f)),
			fadd(fadd(fmul(c.xy, dx), fmul(c.yy, dy)), fmul(c.zy, 1.0f)),
			fadd(fadd(fmul(c.xz, dx), fmul(c.yz, dy)), fmul(c.zz, 1.0f)));
		const float len = __fsqrt_rn(fadd(fadd(fmul(r.x, r.x), fmul(r.y, r.y)), fmul(r.z, r.z)));
		d = v3(fdiv(r.x, len), fdiv(r.y, len), fdiv(r.z, len));
		o = v3(c.px, c.py, c.pz);
	}
	// anti-aliased thin-lens ray (4 RNG draws)
	__device__ __forceinline__ void camera_generate_ray(const DCamera& c, uint32_t x, uint32_t y, Rng& rng, float3& o, float3& d)
	{
		float dx = (((float(x) + 0.5f) / float(c.width)) - 0.5f) * c.tana;
		float dy = (((float(y) + 0.5f) / float(c.height)) - 0.5f) * (-c.tana / c.aspect);
		const float jitter = 0.5f / float(c.width);
		dx += jitter * rng.next_signed();
		dy += jitter * rng.next_signed(); // sic: x resolution on both axes (cuda_camera.cuh:352-355)
		const float3 focal = f3(dx, dy, 1.0f) * c.focal_distance;
		const float angle = rng.next() * 6.2831853f;
		const float radius = sqrtf(rng.next()) * c.aperture;
		float s, cs;
		__sincosf(angle, &s, &cs);
		const float3 lo = f3(radius * s, radius * cs, 0.0f);
		const float3 ld = focal - lo;
		const float3 ax = f3(c.xx, c.xy, c.xz), ay = f3(c.yx, c.yy, c.yz), az = f3(c.zx, c.zy, c.zz);
		o = ax * lo.x + ay * lo.y + az * lo.z + f3(c.px, c.py, c.pz);
		d = normalize(ax * ld.x + ay * ld.y + az * ld.z);
	}

	// ------------------------------------------------------------------ any-hit attenuation (cuda_instance.cuh:105-112)
	__device__ __forceinline__ float4 shadow_attenuation(const DScene& sc, const uint32_t i, const float tb1, const float tb2,
		const uint32_t mat_offset, const uint32_t mat_count)
	{
		const float4 c0 = __ldg(sc.tri_cold + 5 * size_t(i));
		const float4 c1 = __ldg(sc.tri_cold + 5 * size_t(i) + 1);
		const float4 c2 = __ldg(sc.tri_cold + 5 * size_t(i) + 2);
		const float4 c3 = __ldg(sc.tri_cold + 5 * size_t(i) + 3);
		const float4 c4 = __ldg(sc.tri_cold + 5 * size_t(i) + 4);
		const float b3 = 1.0f - tb1 - tb2;
		const float u = c0.w * b3 + c2.w * tb1 + c4.x * tb2;
		const float v = c1.w * b3 + c3.w * tb1 + c4.y * tb2;
		const uint32_t slot = __float_as_uint(__ldg(sc.tri_hot + 3 * size_t(i) + 2).y);
		return material_opacity_color(sc, sc.materials[instance_material(sc, mat_offset, mat_count, slot)], u, v);
	}
}
