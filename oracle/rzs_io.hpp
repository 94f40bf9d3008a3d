// "RZSC" container: a flat list of named raw arrays, used to move flattened scenes, ray sets, hit
// dumps and accumulators between the C++ tools and Python (rayzath_b200/rzs.py reads/writes the same).
// layout: magic "RZSC" | u32 version=1 | u32 n | n x { char name[32] | u32 elem_size | u32 pad | u64 count | bytes, padded to 8 }
#ifndef RZS_IO_HPP
#define RZS_IO_HPP
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rzs
{
	struct Array
	{
		uint32_t elem_size = 1;
		uint64_t count = 0;
		std::vector<uint8_t> bytes;
		template <typename T> const T* as() const { return reinterpret_cast<const T*>(bytes.data()); }
		template <typename T> T* as() { return reinterpret_cast<T*>(bytes.data()); }
	};

	class Writer
	{
		std::vector<std::pair<std::string, Array>> m_arrays;
	public:
		void add(const std::string& name, const void* data, uint32_t elem_size, uint64_t count)
		{
			Array a;
			a.elem_size = elem_size;
			a.count = count;
			a.bytes.resize(size_t(elem_size) * count);
			if (!a.bytes.empty()) std::memcpy(a.bytes.data(), data, a.bytes.size());
			m_arrays.emplace_back(name, std::move(a));
		}
		template <typename T> void add(const std::string& name, const std::vector<T>& v)
		{
			add(name, v.data(), uint32_t(sizeof(T)), v.size());
		}
		template <typename T> void addValue(const std::string& name, const T& v) { add(name, &v, uint32_t(sizeof(T)), 1); }
		void write(const std::string& path) const
		{
			FILE* f = std::fopen(path.c_str(), "wb");
			if (!f) throw std::runtime_error("cannot open " + path);
			const uint32_t version = 1, n = uint32_t(m_arrays.size());
			std::fwrite("RZSC", 1, 4, f);
			std::fwrite(&version, 4, 1, f);
			std::fwrite(&n, 4, 1, f);
			for (const auto& [name, a] : m_arrays)
			{
				char nm[32] = {};
				std::strncpy(nm, name.c_str(), 31);
				const uint32_t pad = 0;
				std::fwrite(nm, 1, 32, f);
				std::fwrite(&a.elem_size, 4, 1, f);
				std::fwrite(&pad, 4, 1, f);
				std::fwrite(&a.count, 8, 1, f);
				if (!a.bytes.empty()) std::fwrite(a.bytes.data(), 1, a.bytes.size(), f);
				const size_t tail = (8 - a.bytes.size() % 8) % 8;
				const char zeros[8] = {};
				if (tail) std::fwrite(zeros, 1, tail, f);
			}
			std::fclose(f);
		}
	};

	inline std::map<std::string, Array> read(const std::string& path)
	{
		FILE* f = std::fopen(path.c_str(), "rb");
		if (!f) throw std::runtime_error("cannot open " + path);
		char magic[4];
		uint32_t version = 0, n = 0;
		if (std::fread(magic, 1, 4, f) != 4 || std::memcmp(magic, "RZSC", 4) != 0) throw std::runtime_error("bad magic in " + path);
		if (std::fread(&version, 4, 1, f) != 1 || std::fread(&n, 4, 1, f) != 1) throw std::runtime_error("short file " + path);
		std::map<std::string, Array> out;
		for (uint32_t i = 0; i < n; ++i)
		{
			char nm[33] = {};
			Array a;
			uint32_t pad;
			if (std::fread(nm, 1, 32, f) != 32 || std::fread(&a.elem_size, 4, 1, f) != 1 ||
				std::fread(&pad, 4, 1, f) != 1 || std::fread(&a.count, 8, 1, f) != 1) throw std::runtime_error("short file " + path);
			a.bytes.resize(size_t(a.elem_size) * a.count);
			if (!a.bytes.empty() && std::fread(a.bytes.data(), 1, a.bytes.size(), f) != a.bytes.size()) throw std::runtime_error("short file " + path);
			const size_t tail = (8 - a.bytes.size() % 8) % 8;
			char zeros[8];
			if (tail && std::fread(zeros, 1, tail, f) != tail) throw std::runtime_error("short file " + path);
			out.emplace(nm, std::move(a));
		}
		std::fclose(f);
		return out;
	}
}
#endif
