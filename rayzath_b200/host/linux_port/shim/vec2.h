// Shim for Math::vec2<T> (un-vendored; see constants.h header note). Semantics follow the CUDA
// mirror type the reference keeps in-tree (cuda_render_parts.cuh:298-495): component-wise
// arithmetic, Rotate(angle) = (x cos + y sin, y cos - x sin).
#ifndef RZ_SHIM_MATH_VEC2_H
#define RZ_SHIM_MATH_VEC2_H
#include <cmath>
#include <cstdint>
#include "angle.h"
namespace Math
{
	template <typename T>
	struct vec2
	{
		T x, y;

		constexpr vec2() noexcept : x(T(0)), y(T(0)) {}
		constexpr explicit vec2(const T v) noexcept : x(v), y(v) {}
		constexpr vec2(const T xx, const T yy) noexcept : x(xx), y(yy) {}
		template <typename U>
		constexpr explicit vec2(const vec2<U>& v) noexcept : x(T(v.x)), y(T(v.y)) {}
		constexpr vec2(const vec2&) = default;
		constexpr vec2& operator=(const vec2&) = default;

		static constexpr T DotProduct(const vec2& a, const vec2& b) noexcept { return a.x * b.x + a.y * b.y; }
		T Magnitude() const { return T(std::sqrt(x * x + y * y)); }
		void Normalize() { const T m = Magnitude(); x /= m; y /= m; }
		vec2 Normalized() const { vec2 v(*this); v.Normalize(); return v; }
		void Rotate(const float angle)
		{
			const float s = sinf(angle), c = cosf(angle);
			const T nx = T(x * c + y * s);
			y = T(y * c - x * s);
			x = nx;
		}
		void Rotate(const angle_radf& angle) { Rotate(angle.value()); }
		vec2 Rotated(const float angle) const { vec2 v(*this); v.Rotate(angle); return v; }
		vec2 Rotated(const angle_radf& angle) const { return Rotated(angle.value()); }

		constexpr vec2 operator-() const noexcept { return vec2(-x, -y); }
		constexpr vec2 operator+(const vec2& v) const noexcept { return vec2(x + v.x, y + v.y); }
		constexpr vec2 operator-(const vec2& v) const noexcept { return vec2(x - v.x, y - v.y); }
		constexpr vec2 operator*(const vec2& v) const noexcept { return vec2(x * v.x, y * v.y); }
		constexpr vec2 operator/(const vec2& v) const { return vec2(x / v.x, y / v.y); }
		constexpr vec2 operator+(const T s) const noexcept { return vec2(x + s, y + s); }  // world.cpp:190
		constexpr vec2 operator-(const T s) const noexcept { return vec2(x - s, y - s); }
		constexpr vec2 operator*(const T s) const noexcept { return vec2(x * s, y * s); }
		constexpr vec2 operator/(const T s) const { return vec2(x / s, y / s); }
		constexpr vec2& operator+=(const vec2& v) noexcept { x += v.x; y += v.y; return *this; }
		constexpr vec2& operator-=(const vec2& v) noexcept { x -= v.x; y -= v.y; return *this; }
		constexpr vec2& operator*=(const vec2& v) noexcept { x *= v.x; y *= v.y; return *this; }
		constexpr vec2& operator/=(const vec2& v) { x /= v.x; y /= v.y; return *this; }
		constexpr vec2& operator*=(const T s) noexcept { x *= s; y *= s; return *this; }
		constexpr vec2& operator/=(const T s) { x /= s; y /= s; return *this; }
		constexpr bool operator==(const vec2& v) const noexcept { return x == v.x && y == v.y; }
		constexpr bool operator!=(const vec2& v) const noexcept { return !(*this == v); }
	};
	template <typename T>
	constexpr vec2<T> operator*(const T s, const vec2<T>& v) noexcept { return v * s; }

	using vec2f = vec2<float>;
	using vec2f32 = vec2<float>;
	using vec2f64 = vec2<double>;
	using vec2i32 = vec2<int32_t>;
	using vec2u32 = vec2<uint32_t>;
	using vec2ui32 = vec2<uint32_t>;
	using vec2u16 = vec2<uint16_t>;
	using vec2ui16 = vec2<uint16_t>;
}
#endif
