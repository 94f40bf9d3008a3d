// Shim for Graphics::Buffer2D<T> / Graphics::Bitmap (un-vendored `Graphics`; see color.h).
// Row-major width*height storage; the method set is the one the reference calls
// (camera.cpp:244-256, loader.cpp:48-50, cpu_engine_renderer.cpp:34, cuda_engine_core.cu:163-240).
#ifndef RZ_SHIM_GRAPHICS_BITMAP_H
#define RZ_SHIM_GRAPHICS_BITMAP_H
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "color.h"
#include "point.h"
namespace Graphics
{
	template <typename T>
	struct Buffer2D
	{
	private:
		uint32_t m_width, m_height;
		std::vector<T> m_data;

	public:
		Buffer2D(const uint32_t width = 1u, const uint32_t height = 1u, const T& value = T{})
			: m_width(std::max(width, 1u)), m_height(std::max(height, 1u))
			, m_data(size_t(m_width) * m_height, value) {}

		uint32_t GetWidth() const noexcept { return m_width; }
		uint32_t GetHeight() const noexcept { return m_height; }
		T* GetMapAddress() noexcept { return m_data.data(); }
		const T* GetMapAddress() const noexcept { return m_data.data(); }

		T& Value(const uint32_t x, const uint32_t y) { return m_data[size_t(y) * m_width + x]; }
		const T& Value(const uint32_t x, const uint32_t y) const { return m_data[size_t(y) * m_width + x]; }
		void SetValue(const uint32_t x, const uint32_t y, const T& v) { Value(x, y) = v; }

		void Resize(const uint32_t width, const uint32_t height)
		{
			const uint32_t w = std::max(width, 1u), h = std::max(height, 1u);
			if (w == m_width && h == m_height) return;
			std::vector<T> data(size_t(w) * h, T{});
			for (uint32_t y = 0; y < std::min(h, m_height); ++y)
				for (uint32_t x = 0; x < std::min(w, m_width); ++x)
					data[size_t(y) * w + x] = m_data[size_t(y) * m_width + x];
			m_data.swap(data);
			m_width = w; m_height = h;
		}
		void Clear(const T& value = T{}) { std::fill(m_data.begin(), m_data.end(), value); }
		// copy `bytes` of raw T elements from `src` into the buffer starting at pixel (x, y) in row-major order
		void CopyFromMemory(const void* src, const size_t bytes, const uint32_t x = 0u, const uint32_t y = 0u)
		{
			const size_t offset = size_t(y) * m_width + x;
			if (offset >= m_data.size()) return;
			const size_t room = (m_data.size() - offset) * sizeof(T);
			std::memcpy(reinterpret_cast<char*>(m_data.data() + offset), src, std::min(bytes, room));
		}
	};
	using Bitmap = Buffer2D<Color>;
}
#endif
