"""gatling_b200 -- B200-native wavefront path-tracing core behind gatling's `gi` renderer API.

  gatling_b200.gi       Python mirror of the reference's gi* API (src/gi/gtl/gi/Gi.h) over the C-ABI
  gatling_b200.scenes   USD-free scene descriptions of the BASELINE.json configs
  gatling_b200.build    in-tree nvcc build of gatling_b200/libgtlc.so (sm_100a)
"""
__version__ = "0.1.0"
