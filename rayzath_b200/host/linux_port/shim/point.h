// Shim for Graphics::Point<T> (un-vendored `Graphics`; see color.h). Call site: cuda_engine_core.cu:204.
#ifndef RZ_SHIM_GRAPHICS_POINT_H
#define RZ_SHIM_GRAPHICS_POINT_H
namespace Graphics
{
	template <typename T>
	struct Point
	{
		T x, y;
		constexpr Point(const T xx = T(0), const T yy = T(0)) noexcept : x(xx), y(yy) {}
	};
}
#endif
