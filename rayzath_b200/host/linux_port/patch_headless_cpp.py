#!/usr/bin/env python3
"""Generate g++-compatible, extended copies of the reference's Application/headless.cpp and headless.hpp at BUILD
time (never committed; the reference sources are read where they lie).

1. LINUX PORTABILITY. headless.cpp:72 brace-initialises a nlohmann::json from another json
   (`const auto json{json_t::parse(...)}`): MSVC copy-constructs, g++ picks the initializer_list constructor and
   wraps the document in a one-element array, after which `json.contains("tasks")` fails. The copy uses `=`.

2. TASK FORMAT. The reference's task entry knows "scene path", "engine", "rpp", "timeout"
   (headless.cpp:56-134); RenderTask::max_depth exists (headless.hpp:16) and is applied (headless.cpp:182) but is never
   read from the file. The copy reads these additional, optional keys:
     "max depth"    1..255 -> RenderTask::max_depth (BASELINE config 1 is depth 8)
     "seed"         u64    -> the engine's RNG seed        (rzb200_engine_configure)
     "devices"      "0,1"  -> CUDA ordinals, one disjoint sample stream per device
     "spp"          float  -> stop when the mean of completed paths per pixel reaches it (rzb200_engine_mean_spp),
                              in addition to the "rpp" pass budget and the timeout
     "accumulator"  path   -> after the task, dump camera 0's float accumulator (width*height*4 f32, rgb sum + sample
                              count) relative to the report directory (rzb200_engine_read_accum)
   The three rzb200_engine_* hooks are declared WEAK: the B200 drop-in engine (cuda_engine_b200.cpp) defines them; in
   a build with the reference's own engines (the test-side reference tool) they are absent and the keys beyond "max depth"
   are ignored -- the same generated file serves both.

usage: patch_headless_cpp.py <reference headless.cpp> <output headless.cpp>
       (headless.hpp is read next to the input and written next to the output)
"""
import os
import sys


def replace_once(text, old, new, what):
    if text.count(old) != 1:
        sys.stderr.write("patch_headless_cpp: pattern not found (%s)\n" % what)
        sys.exit(1)
    return text.replace(old, new)


HOOKS = '''
// ---- added by linux_port/patch_headless_cpp.py: optional engine hooks (weak: absent with the reference's own engines)
extern "C"
{
	void rzb200_engine_configure(int has_seed, unsigned long long seed, const char* devices) __attribute__((weak));
	double rzb200_engine_mean_spp(unsigned camera) __attribute__((weak));
	int rzb200_engine_read_accum(unsigned camera, float* rgba_f32, size_t floats) __attribute__((weak));
}
'''

READ_KEYS = '''
				// ---- added keys (patch_headless_cpp.py)
				if (entry_json.contains("max depth"))
					task.max_depth = static_cast<uint8_t>(std::clamp<int>(static_cast<int>(entry_json["max depth"]), 1, 255));
				if (entry_json.contains("seed"))
				{
					task.has_seed = true;
					task.seed = static_cast<uint64_t>(entry_json["seed"]);
				}
				if (entry_json.contains("devices"))
					task.devices = static_cast<std::string>(entry_json["devices"]);
				if (entry_json.contains("spp"))
					task.spp = static_cast<float>(entry_json["spp"]);
				if (entry_json.contains("accumulator"))
					task.accumulator = static_cast<std::string>(entry_json["accumulator"]);

				return task;
			};
'''


def main() -> int:
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src, encoding="utf-8", errors="replace").read()
    text = replace_once(text, "const auto json{json_t::parse(file, nullptr, true, true)};",
                        "const auto json = json_t::parse(file, nullptr, true, true);", "json brace init")
    text = replace_once(text, "namespace RayZath::Headless\n{", HOOKS + "\nnamespace RayZath::Headless\n{", "namespace")
    text = replace_once(text, "\n				return task;\n			};\n", READ_KEYS, "createTask return")
    # seed / devices reach the engine once the scene is loaded and the engine exists
    text = replace_once(
        text, "		engine.renderConfig().tracing().maxDepth(task.max_depth);\n",
        "		engine.renderConfig().tracing().maxDepth(task.max_depth);\n"
        "		if (rzb200_engine_configure && (task.has_seed || !task.devices.empty()))\n"
        "			rzb200_engine_configure(task.has_seed ? 1 : 0, task.seed, task.devices.c_str());\n", "maxDepth")
    # "spp" stop next to the timeout
    text = replace_once(
        text, "					if (task_duration.count() >= task.timeout)\n						break;\n",
        "					if (task_duration.count() >= task.timeout)\n						break;\n"
        "					if (task.spp > 0.0f && rzb200_engine_mean_spp && rzb200_engine_mean_spp(0) >= double(task.spp))\n"
        "						break;\n", "timeout break")
    # accumulator dump after the render loop of an engine
    text = replace_once(
        text, "				result.duration = duration;\n",
        "				result.duration = duration;\n"
        "				if (!task.accumulator.empty() && rzb200_engine_read_accum && cameras.count() && cameras[0])\n"
        "				{\n"
        "					std::vector<float> acc(size_t(cameras[0]->width()) * cameras[0]->height() * 4);\n"
        "					if (rzb200_engine_read_accum(0, acc.data(), acc.size()) == 0)\n"
        "					{\n"
        "						std::ofstream dump(report_dir / task.accumulator, std::ios::binary);\n"
        "						dump.write(reinterpret_cast<const char*>(acc.data()), std::streamsize(acc.size() * sizeof(float)));\n"
        "						std::cout << \"Saved accumulator of \\\"\" << cameras[0]->name() << \"\\\" to \" << task.accumulator << \"\\n\";\n"
        "					}\n"
        "				}\n", "duration")
    open(dst, "w", encoding="utf-8").write(text)

    hsrc = os.path.join(os.path.dirname(src), "headless.hpp")
    hdst = os.path.join(os.path.dirname(dst), "headless.hpp")
    h = open(hsrc, encoding="utf-8", errors="replace").read()
    h = replace_once(h, "		uint8_t max_depth = 16;\n	};\n	struct TaskResult",
                     "		uint8_t max_depth = 16;\n"
                     "		// added by linux_port/patch_headless_cpp.py (optional task keys)\n"
                     "		bool has_seed = false;\n		uint64_t seed = 0;\n		std::string devices;\n"
                     "		float spp = 0.0f;\n		std::string accumulator;\n	};\n	struct TaskResult", "RenderTask")
    open(hdst, "w", encoding="utf-8").write(h)
    return 0


if __name__ == "__main__":
    sys.exit(main())
