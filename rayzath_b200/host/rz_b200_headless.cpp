// Headless entry of the drop-in build: the reference's own `--headless <tasks.json> [report_dir] [-r]` runner
// (Application/headless.cpp, compiled in place) linked against the reference host library and THIS repo's
// RayZath::Cuda::Engine (cuda_engine_b200.cpp -> librzb200.so). It is Application/main.cpp:41-77 without the
// Windows CRT headers (main.cpp:1-6, 21-22) and without the GUI branch.
#include <filesystem>
#include <iostream>
#include <stdexcept>

#include "args.hpp"
#include "rzexception.hpp"
#include "headless.hpp"

int main(int argc, char* argv[])
{
	try
	{
		auto arg_def = RayZath::Args{}
			.arg(RayZath::Args::Arg({"-h", "--help"}, "Prints help message.", {}))
			.arg(RayZath::Args::Arg({"--headless"}, "Execute rendering tasks without UI and generate a report.",
				{RayZath::Args::Option("task_path", true), RayZath::Args::Option("report_path", false)}))
			.arg(RayZath::Args::Arg({"-r", "--render"}, "When specified --headless, also saves rendered images.", {}));
		auto args{arg_def.parse(argc - 1, argv + 1)};
		if (args.contains("-h") || args.contains("--help") || !args.contains("--headless"))
		{
			std::cout << arg_def.usageString() << std::endl;
			return args.contains("--headless") ? 0 : 2;
		}
		const auto& params = args["--headless"];
		std::filesystem::path task_path{}, report_path{};
		if (params.size() > 0) task_path.assign(params[0]);
		if (params.size() > 1) report_path.assign(params[1]);
		return RayZath::Headless::Headless::instance().run(task_path, report_path, args.contains("-r"));
	}
	catch (std::exception& ex)
	{
		std::cerr << ex.what() << '\n';
		return 1;
	}
}
