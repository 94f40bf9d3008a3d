// Shim for the un-vendored `Math` library (Greketrotny/Math, no pinned version).
// LINUX PORTABILITY LAYER: lets the reference's own host sources compile on Linux (drop-in engine build, oracle build).
// Written from the call sites in /root/reference/RayZath (e.g. world.cpp:179, camera.hpp:130);
// "parity unpinned" at this level: the literal used upstream is unknown, (float)pi is assumed.
#ifndef RZ_SHIM_MATH_CONSTANTS_H
#define RZ_SHIM_MATH_CONSTANTS_H
namespace Math
{
	template <typename T>
	struct constants
	{
		static constexpr T pi = T(3.14159265358979323846);
		static constexpr T r_pi = T(1.0 / 3.14159265358979323846);
	};
}
#endif
