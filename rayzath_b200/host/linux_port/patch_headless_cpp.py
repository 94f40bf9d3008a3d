#!/usr/bin/env python3
"""Generate a g++-compatible copy of the reference's Application/headless.cpp at BUILD time (never committed).

LINUX PORTABILITY LAYER. headless.cpp:72 brace-initialises a nlohmann::json from another json
(`const auto json{json_t::parse(...)}`): MSVC copy-constructs, g++ picks the initializer_list constructor and
wraps the document in a one-element array, after which `json.contains("tasks")` fails. The generated copy uses
`=` initialisation; nothing else is touched.

usage: patch_headless_cpp.py <reference headless.cpp> <output path>
"""
import sys


def main() -> int:
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src, encoding="utf-8", errors="replace").read()
    old = "const auto json{json_t::parse(file, nullptr, true, true)};"
    if text.count(old) != 1:
        sys.stderr.write("patch_headless_cpp: pattern not found\n")
        return 1
    open(dst, "w", encoding="utf-8").write(text.replace(old, "const auto json = json_t::parse(file, nullptr, true, true);"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
