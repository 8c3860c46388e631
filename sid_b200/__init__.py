"""sid_b200: B200-native implementation of EvolBioInf/sid's per-site genotype-calling path.

The package is a thin host mirror of the reference interface (call.hpp) over libsidgpu.so
(hand-written sm_100a kernels behind the C ABI of include/sidgpu.h)."""
from .api import (CSV_HEADER, METHODS, Context, MalformedPileup, OutputRecord, SidGpuError, callBayes,  # noqa: F401
                  callLikelihoodRatio, callQualityBasedSimple, callSiteMLError, call_columns, columns_to_arrow, parse_csv_rows,
                  sid_csv)
