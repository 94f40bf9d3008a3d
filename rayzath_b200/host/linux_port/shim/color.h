// Shim for Graphics::Color / Graphics::ColorF (un-vendored `Graphics` library, no pinned version;
// SURVEY.md §8c). LINUX PORTABILITY LAYER. Semantics follow the CUDA mirror types the reference
// keeps in-tree (cuda_render_parts.cuh:497-683 ColorF, :685-825 ColorU): ColorF(Color) = /255.0f,
// Blend(a,b,t) = a + (b-a)*t, arithmetic is 4-channel component-wise.
// Unknowable from the tree ("parity unpinned"): Palette byte values other than White/Black.
#ifndef RZ_SHIM_GRAPHICS_COLOR_H
#define RZ_SHIM_GRAPHICS_COLOR_H
#include <cstdint>
namespace Graphics
{
	struct Color
	{
		uint8_t red, green, blue, alpha;

		constexpr Color() noexcept : red(0), green(0), blue(0), alpha(0xFF) {}
		constexpr Color(const uint8_t r, const uint8_t g, const uint8_t b, const uint8_t a = 0xFF) noexcept
			: red(r), green(g), blue(b), alpha(a) {}
		constexpr explicit Color(const uint8_t v) noexcept : red(v), green(v), blue(v), alpha(0xFF) {}

		constexpr bool operator==(const Color& o) const noexcept
		{
			return red == o.red && green == o.green && blue == o.blue && alpha == o.alpha;
		}
		constexpr bool operator!=(const Color& o) const noexcept { return !(*this == o); }

		struct Palette
		{
			static const Color White, Black, Grey, LightGrey, DarkGrey, Red, Green, Blue;
		};
	};
	inline constexpr Color Color::Palette::White{0xFF, 0xFF, 0xFF, 0xFF};
	inline constexpr Color Color::Palette::Black{0x00, 0x00, 0x00, 0xFF};
	inline constexpr Color Color::Palette::Grey{0x80, 0x80, 0x80, 0xFF};
	inline constexpr Color Color::Palette::LightGrey{0xC0, 0xC0, 0xC0, 0xFF};
	inline constexpr Color Color::Palette::DarkGrey{0x40, 0x40, 0x40, 0xFF};
	inline constexpr Color Color::Palette::Red{0xFF, 0x00, 0x00, 0xFF};
	inline constexpr Color Color::Palette::Green{0x00, 0xFF, 0x00, 0xFF};
	inline constexpr Color Color::Palette::Blue{0x00, 0x00, 0xFF, 0xFF};

	struct ColorF
	{
		float red, green, blue, alpha;

		constexpr ColorF() noexcept : red(1.0f), green(1.0f), blue(1.0f), alpha(1.0f) {}
		constexpr explicit ColorF(const float v) noexcept : red(v), green(v), blue(v), alpha(v) {}
		constexpr ColorF(const float r, const float g, const float b, const float a = 1.0f) noexcept
			: red(r), green(g), blue(b), alpha(a) {}
		constexpr explicit ColorF(const Color& c) noexcept
			: red(c.red / 255.0f), green(c.green / 255.0f), blue(c.blue / 255.0f), alpha(c.alpha / 255.0f) {}

		constexpr ColorF operator+(const ColorF& o) const noexcept
		{
			return ColorF(red + o.red, green + o.green, blue + o.blue, alpha + o.alpha);
		}
		constexpr ColorF operator-(const ColorF& o) const noexcept
		{
			return ColorF(red - o.red, green - o.green, blue - o.blue, alpha - o.alpha);
		}
		constexpr ColorF operator*(const ColorF& o) const noexcept
		{
			return ColorF(red * o.red, green * o.green, blue * o.blue, alpha * o.alpha);
		}
		constexpr ColorF operator/(const ColorF& o) const
		{
			return ColorF(red / o.red, green / o.green, blue / o.blue, alpha / o.alpha);
		}
		constexpr ColorF operator*(const float f) const noexcept
		{
			return ColorF(red * f, green * f, blue * f, alpha * f);
		}
		constexpr ColorF operator/(const float f) const
		{
			return ColorF(red / f, green / f, blue / f, alpha / f);
		}
		constexpr ColorF& operator+=(const ColorF& o) noexcept { return *this = *this + o; }
		constexpr ColorF& operator-=(const ColorF& o) noexcept { return *this = *this - o; }
		constexpr ColorF& operator*=(const ColorF& o) noexcept { return *this = *this * o; }
		constexpr ColorF& operator/=(const ColorF& o) { return *this = *this / o; }
		constexpr ColorF& operator*=(const float f) noexcept { return *this = *this * f; }
		constexpr ColorF& operator/=(const float f) { return *this = *this / f; }

		static constexpr ColorF Blend(const ColorF& a, const ColorF& b, const float t) noexcept
		{
			return a + (b - a) * t;
		}
		constexpr void Blend(const ColorF& c, const float t) noexcept { *this = Blend(*this, c, t); }
	};
}
#endif
