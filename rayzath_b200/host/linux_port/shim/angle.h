// Shim for Math::angle (un-vendored; see constants.h header note). Call sites: camera.hpp:38,130,
// render_parts.hpp:123,169,198,237. Only radians-in-float with value() is used by the reference.
#ifndef RZ_SHIM_MATH_ANGLE_H
#define RZ_SHIM_MATH_ANGLE_H
#include "constants.h"
namespace Math
{
	enum class angle_unit { rad, deg };
	template <angle_unit U, typename T = float>
	struct angle
	{
	private:
		T m_value;
	public:
		constexpr angle(const T v = T(0)) noexcept : m_value(v) {}
		constexpr const T& value() const noexcept { return m_value; }
		constexpr T& value() noexcept { return m_value; }  // camera.cpp:110 assigns through value()
		// json_loader.cpp:141,687 assign a json value (implicitly convertible to float) to an angle
		constexpr angle& operator=(const T v) noexcept { m_value = v; return *this; }
		constexpr angle operator-() const noexcept { return angle(-m_value); }
		constexpr angle operator+(const angle& o) const noexcept { return angle(m_value + o.m_value); }
		constexpr angle operator-(const angle& o) const noexcept { return angle(m_value - o.m_value); }
		constexpr angle operator*(const T s) const noexcept { return angle(m_value * s); }
		constexpr angle operator/(const T s) const noexcept { return angle(m_value / s); }
		constexpr angle& operator+=(const angle& o) noexcept { m_value += o.m_value; return *this; }
		constexpr angle& operator-=(const angle& o) noexcept { m_value -= o.m_value; return *this; }
		constexpr bool operator==(const angle& o) const noexcept { return m_value == o.m_value; }
		constexpr bool operator!=(const angle& o) const noexcept { return m_value != o.m_value; }
		constexpr bool operator<(const angle& o) const noexcept { return m_value < o.m_value; }
		constexpr bool operator>(const angle& o) const noexcept { return m_value > o.m_value; }
	};
	using angle_radf = angle<angle_unit::rad, float>;
	using angle_degf = angle<angle_unit::deg, float>;
}
#endif
