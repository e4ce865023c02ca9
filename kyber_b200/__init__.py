"""kyber_b200 -- B200-native batch group arithmetic behind dedis/kyber's Group/Point/Scalar and
pairing.Suite interfaces.  The product is kyber_b200/libb2kyber.so (hand-written sm_100a CUDA behind
the C ABI of include/b2kyber.h); this package is the thin Python host side used by tests and bench."""
from .capi import Engine, Comm, B2KError, load_library, LIB_PATH  # noqa: F401
