#!/usr/bin/env python3
"""Generate a g++-compatible copy of the reference's world.hpp at BUILD time (never committed).

LINUX PORTABILITY LAYER (used by the drop-in host build and by the oracle build). g++ 13 rejects explicit specialisations at class scope
(/root/reference/RayZath/world.hpp:143-194, `template<> struct CommonMeshParameters<...>` inside
`class World`); MSVC accepts them. Partial specialisations ARE legal at class scope, so the
generated header adds a defaulted dummy parameter and turns each full specialisation into a
partial one. Nothing else in the file is touched. The output keeps the original include guard
(WORLD_H) and is force-included (-include) so the in-place `#include "world.hpp"` of every
reference source becomes a no-op.

usage: patch_world_hpp.py <reference world.hpp> <output path>
"""
import re
import sys


def main() -> int:
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src, encoding="utf-8", errors="replace").read()
    n_primary = len(re.findall(r"template\s*<CommonMesh T>\s*struct CommonMeshParameters\s*\{\s*\}\s*;", text))
    text = re.sub(
        r"template\s*<CommonMesh T>(\s*)struct CommonMeshParameters\s*\{\s*\}\s*;",
        r"template <CommonMesh T, typename RzDummy = void>\1struct CommonMeshParameters {};",
        text)
    text, n_spec = re.subn(
        r"template\s*<\s*>(\s*)struct CommonMeshParameters<\s*(CommonMesh::\w+)\s*>",
        r"template <typename RzDummy>\1struct CommonMeshParameters<\2, RzDummy>",
        text)
    if n_primary != 1 or n_spec == 0:
        sys.stderr.write("patch_world_hpp: pattern not found (primary=%d, specialisations=%d)\n" % (n_primary, n_spec))
        return 1
    open(dst, "w", encoding="utf-8").write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
