"""trust4_b200 -- B200-native stage-1 assembly hot path of TRUST4 (host-side Python mirror).

The product is the CUDA engine in csrc/ behind the C ABI of include/trust4_b200.h;
this package loads it with ctypes (trust4_b200.api) and provides the synthetic
workload generator (trust4_b200.synth).  There is no CPU fallback.
"""
__version__ = "0.1.0"
