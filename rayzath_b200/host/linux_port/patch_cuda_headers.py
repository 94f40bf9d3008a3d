#!/usr/bin/env python3
"""Generate nvcc-12.9-compatible copies of two reference CUDA headers at BUILD time (never committed).

LINUX PORTABILITY LAYER (oracle build of the reference's own CUDA engine, used as a GPU comparand). MSVC's front
end accepts what nvcc's does not (SURVEY.md 8c blocker 5):
  cuda_render_parts.cuh:716-719   functional casts `unsigned char(expr)`            -> `(unsigned char)(expr)`
  cuda_buffer.cuh:147,182-194,... dependent type without `typename` in template arguments
                                  `CudaVectorType<T>::type`                          -> `typename CudaVectorType<T>::type`
Nothing else is touched. The outputs keep the original include guards and are force-included (-include), so the
in-place `#include "cuda_render_parts.cuh"` of every reference source becomes a no-op.

usage: patch_cuda_headers.py <reference RayZath dir> <output dir>
"""
import os
import re
import sys


def main() -> int:
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(dst, exist_ok=True)
    t = open(os.path.join(src, "cuda_render_parts.cuh"), encoding="utf-8", errors="replace").read()
    n1 = t.count("unsigned char(")
    open(os.path.join(dst, "cuda_render_parts.cuh"), "w", encoding="utf-8").write(t.replace("unsigned char(", "(unsigned char)("))
    t = open(os.path.join(src, "cuda_buffer.cuh"), encoding="utf-8", errors="replace").read()
    pat = r"(?<!typename )CudaVectorType<T>::type"
    n2 = len(re.findall(pat, t))
    open(os.path.join(dst, "cuda_buffer.cuh"), "w", encoding="utf-8").write(re.sub(pat, "typename CudaVectorType<T>::type", t))
    if n1 == 0 or n2 == 0:
        sys.stderr.write("patch_cuda_headers: patterns not found (%d, %d)\n" % (n1, n2))
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
