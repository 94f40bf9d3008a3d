"""rayzath_b200: B200-native (sm_100a) wavefront path tracer behind RayZath's CUDA-engine boundary.

The product is rayzath_b200/librzb200.so (C ABI, include/rzb200.h) and the C++ drop-in engine in
rayzath_b200/host/. The Python modules are the host-side harness used by tests and bench.py:
  capi     ctypes binding of the C ABI
  world    scene graph mirror -> flattened C-ABI arrays / reference scene files
  scenes   the procedural scenes of BASELINE.json's configurations
  parallel one-process-per-GPU sharding of sample streams and the NCCL accumulator reduce
  rzs      array container shared with the C++ tools
"""
__version__ = "0.1.0"
