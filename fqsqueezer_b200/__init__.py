"""fqsqueezer_b200 -- B200-native k-mer statistics engine for FQSqueezer (hot path only).

The product path is libfqsk.so (hand-written sm_100a CUDA behind the C-ABI of include/fqsk.h); this package is the
thin Python host mirror used by tests and bench.py.  It never imports the CPU oracle and has no CPU fallback:
creating an engine without the library or without a CUDA device raises."""
from .engine import FqskError, KmerEngine, kmer_params, load_library  # noqa: F401
