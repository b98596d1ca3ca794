"""hisstools_library_b200 -- B200-native partitioned convolution behind the HISSTools_Library API.

The package holds the CUDA library (csrc/, built in-tree into lib/libhisstools_b200.so), its ctypes
binding and the host-side mirror of the reference classes for this path.  Importing it does not
need a GPU; calling any transform or convolver does, and raises HissError otherwise.
"""
from ._abi import HissError, lib as load_library          # noqa: F401
from .errors import *                                      # noqa: F401,F403
from .errors import ConvolveError, LatencyMode             # noqa: F401
from .fft import (Split, Setup, hisstools_create_setup, hisstools_destroy_setup, hisstools_fft,      # noqa: F401
                  hisstools_ifft, hisstools_rfft, hisstools_rifft, hisstools_zip, hisstools_unzip,
                  hisstools_unzip_zero)
from .convolve import PartitionedConvolve, MonoConvolve, NToMonoConvolve, Convolver, partition_scheme   # noqa: F401
from .spectral import spectral_processor, EdgeMode                                                   # noqa: F401
from .audiofile import IAudioFile, OAudioFile                                                                  # noqa: F401
