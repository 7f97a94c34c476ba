"""fqgrep_b200 -- B200-native (sm_100a) implementation of fqgrep's read-matching hot path.

The product is libfqgrep_b200.so (C ABI, include/fqgrep_b200.h); this package is the thin host-side mirror of
the reference's Matcher / MatcherFactory API over it.  There is no CPU matching path.
"""
from .matcher import (BITMASK, BITMASK_SET, FIXED, FIXED_SET, INTERLEAVED, NAMES, PAIRED, REGEX, REGEX_SET, SINGLE, FqgrepError, Matcher,  # noqa: F401
                      MatcherFactory, MatcherOpts, FastqStream, fastq_cut, context, expand_iupac, expand_iupac_fixed_pattern, expand_iupac_regex, is_fastq_path, is_gzip_path, read_input,
                      validate_fixed_pattern,
                      write_selected)
from .driver import Opts, fqgrep, fqgrep_from_opts, read_patterns, read_query_names  # noqa: E402,F401
