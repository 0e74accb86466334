"""dna_factory_b200 -- B200-native population-generation hot path of ochrzan/dna-factory.

Layout: csrc/ (CUDA kernels + C ABI, built to _lib/libdnaf_b200.so), _native.py (ctypes binding),
host.py (flattening of the reference's SNP / sample objects), pop_factory.py (the reference's CLI and
PopulationFactory surface re-hosted on the GPU path).
"""
__version__ = "0.1.0"
