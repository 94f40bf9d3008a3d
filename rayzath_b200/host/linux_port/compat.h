// Force-included (-include) when compiling the reference's own sources on Linux for the oracle.
// LINUX PORTABILITY LAYER for building the reference host sources in place (drop-in engine, oracle). The reference is written against MSVC's <cmath>, which puts the
// float-suffixed C functions in namespace std (std::tanf, std::sqrtf ...; e.g.
// cpu_engine_kernel.cpp:186,214,235-238,543,640,717,837,847,864); libstdc++ 13 does not.
#ifndef RZ_ORACLE_COMPAT_H
#define RZ_ORACLE_COMPAT_H
#ifdef __cplusplus
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <string>
#include <memory>
#include <limits>
#include <stdexcept>
#include <algorithm>
#include <functional>
#include <atomic>
#include <mutex>
#include <thread>
#include <condition_variable>
#include <chrono>
#include <map>
#include <vector>
#include <array>
#include <iostream>
#include <sstream>
#include <fstream>
namespace std
{
	using ::tanf; using ::sinf; using ::cosf; using ::sqrtf; using ::fmodf; using ::powf;
	using ::logf; using ::expf; using ::acosf; using ::asinf; using ::atanf; using ::atan2f;
	using ::fabsf; using ::floorf; using ::ceilf; using ::log2f; using ::log10f; using ::truncf;
}

// json_loader.cpp:470 concatenates a string literal with a json value (MSVC resolves it through
// the implicit json -> std::string conversion); give ADL an exact match instead.
#include "lib/Json/json.hpp"
namespace nlohmann
{
	inline std::string operator+(const char* lhs, const json& rhs)
	{
		return std::string(lhs) + (rhs.is_string() ? rhs.get<std::string>() : rhs.dump());
	}
}
#endif
#endif
