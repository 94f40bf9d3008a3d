// Shim for Math::vec3<T> (un-vendored; see constants.h header note). Semantics follow the CUDA
// mirror type the reference keeps in-tree (cuda_render_parts.cuh:15-297): dot = x*x'+y*y'+z*z'
// left to right, cross = (y z' - z y', z x' - x z', x y' - y x'), rotations with the same
// sign conventions (:116-182), operator/ = true per-component division.
// Unknowable from the tree ("parity unpinned"): Normalize() is restated as division by
// sqrtf(x^2+y^2+z^2); Similarity as dot/(|a||b|).
#ifndef RZ_SHIM_MATH_VEC3_H
#define RZ_SHIM_MATH_VEC3_H
#include <cmath>
#include <cstdint>
#include "vec2.h"
namespace Math
{
	template <typename T>
	struct vec3
	{
		T x, y, z;

		constexpr vec3() noexcept : x(T(0)), y(T(0)), z(T(0)) {}
		constexpr explicit vec3(const T v) noexcept : x(v), y(v), z(v) {}
		constexpr vec3(const T xx, const T yy, const T zz) noexcept : x(xx), y(yy), z(zz) {}
		template <typename U>
		constexpr explicit vec3(const vec3<U>& v) noexcept : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}
		constexpr vec3(const vec3&) = default;
		constexpr vec3& operator=(const vec3&) = default;

		static constexpr T DotProduct(const vec3& a, const vec3& b) noexcept
		{
			return a.x * b.x + a.y * b.y + a.z * b.z;
		}
		static constexpr vec3 CrossProduct(const vec3& a, const vec3& b) noexcept
		{
			return vec3(
				a.y * b.z - a.z * b.y,
				a.z * b.x - a.x * b.z,
				a.x * b.y - a.y * b.x);
		}
		static T Similarity(const vec3& a, const vec3& b)
		{
			return DotProduct(a, b) / (a.Magnitude() * b.Magnitude());
		}
		static T Distance(const vec3& a, const vec3& b) { return (a - b).Magnitude(); }

		T Magnitude() const { return T(sqrtf(float(x * x + y * y + z * z))); }
		constexpr T MagnitudeSquared() const noexcept { return x * x + y * y + z * z; }
		void Normalize()
		{
			const T m = Magnitude();
			x /= m; y /= m; z /= m;
		}
		vec3 Normalized() const { vec3 v(*this); v.Normalize(); return v; }
		constexpr void Reverse() noexcept { x = -x; y = -y; z = -z; }
		constexpr vec3 Reversed() const noexcept { return vec3(-x, -y, -z); }

		void RotateX(const float angle)
		{
			const float s = sinf(angle), c = cosf(angle);
			const T ny = T(y * c + z * s);
			z = T(y * -s + z * c);
			y = ny;
		}
		void RotateY(const float angle)
		{
			const float s = sinf(angle), c = cosf(angle);
			const T nx = T(x * c - z * s);
			z = T(x * s + z * c);
			x = nx;
		}
		void RotateZ(const float angle)
		{
			const float s = sinf(angle), c = cosf(angle);
			const T nx = T(x * c + y * s);
			y = T(x * -s + y * c);
			x = nx;
		}
		void RotateXYZ(const vec3& rot) { RotateX(rot.x); RotateY(rot.y); RotateZ(rot.z); }
		void RotateZYX(const vec3& rot) { RotateZ(rot.z); RotateY(rot.y); RotateX(rot.x); }
		vec3 RotatedX(const float a) const { vec3 v(*this); v.RotateX(a); return v; }
		vec3 RotatedY(const float a) const { vec3 v(*this); v.RotateY(a); return v; }
		vec3 RotatedZ(const float a) const { vec3 v(*this); v.RotateZ(a); return v; }
		vec3 RotatedXYZ(const vec3& r) const { vec3 v(*this); v.RotateXYZ(r); return v; }
		vec3 RotatedZYX(const vec3& r) const { vec3 v(*this); v.RotateZYX(r); return v; }

		constexpr vec3 operator-() const noexcept { return vec3(-x, -y, -z); }
		constexpr vec3 operator+(const vec3& v) const noexcept { return vec3(x + v.x, y + v.y, z + v.z); }
		constexpr vec3 operator-(const vec3& v) const noexcept { return vec3(x - v.x, y - v.y, z - v.z); }
		constexpr vec3 operator*(const vec3& v) const noexcept { return vec3(x * v.x, y * v.y, z * v.z); }
		constexpr vec3 operator/(const vec3& v) const { return vec3(x / v.x, y / v.y, z / v.z); }
		constexpr vec3 operator+(const T s) const noexcept { return vec3(x + s, y + s, z + s); }
		constexpr vec3 operator-(const T s) const noexcept { return vec3(x - s, y - s, z - s); }
		constexpr vec3 operator*(const T s) const noexcept { return vec3(x * s, y * s, z * s); }
		constexpr vec3 operator/(const T s) const { return vec3(x / s, y / s, z / s); }
		constexpr vec3& operator+=(const vec3& v) noexcept { x += v.x; y += v.y; z += v.z; return *this; }
		constexpr vec3& operator-=(const vec3& v) noexcept { x -= v.x; y -= v.y; z -= v.z; return *this; }
		constexpr vec3& operator*=(const vec3& v) noexcept { x *= v.x; y *= v.y; z *= v.z; return *this; }
		constexpr vec3& operator/=(const vec3& v) { x /= v.x; y /= v.y; z /= v.z; return *this; }
		constexpr vec3& operator*=(const T s) noexcept { x *= s; y *= s; z *= s; return *this; }
		constexpr vec3& operator/=(const T s) { x /= s; y /= s; z /= s; return *this; }
		constexpr bool operator==(const vec3& v) const noexcept { return x == v.x && y == v.y && z == v.z; }
		constexpr bool operator!=(const vec3& v) const noexcept { return !(*this == v); }
	};
	template <typename T>
	constexpr vec3<T> operator*(const T s, const vec3<T>& v) noexcept { return v * s; }

	using vec3f = vec3<float>;
	using vec3f32 = vec3<float>;
	using vec3f64 = vec3<double>;
	using vec3i32 = vec3<int32_t>;
	using vec3u32 = vec3<uint32_t>;
	using vec3ui32 = vec3<uint32_t>;
}
#endif
